"""profiles/r02_parity_report.jsonl (written by `pytest -m gpu`, tests/helpers.py: parity) -> a markdown summary."""
import collections
import json
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "profiles/r02_parity_report.jsonl"
recs = [json.loads(l) for l in open(src)]
cmp_ = [r for r in recs if "cuda_fro" in r]
groups = collections.OrderedDict()
for r in cmp_:
    key = r["what"].split(":")[0]
    groups.setdefault(key, []).append(r)
print("# Round-2 GPU parity report (B200, `pytest tests -m gpu`)\n")
print("Criterion per tensor (tests/helpers.py): `fro(CUDA, fp32 oracle) <= 1.5 x max(fro(eager-bf16 oracle, fp32 oracle), 1.5e-3)` and")
print("`max(CUDA) <= 3 x max(max(eager), 4e-3)`; fro = relative Frobenius error, max = max abs error / max abs reference.\n")
print(f"{len(cmp_)} tensor comparisons, {sum(1 for r in cmp_ if not r['ok'])} outside the criterion.\n")
print("| case | tensors | worst cuda fro | eager fro there | worst ratio cuda/eager (fro) | median ratio | worst ratio (max-norm) |")
print("|---|---|---|---|---|---|---|")
for k, rs in groups.items():
    ratio = [r["cuda_fro"] / max(r["eager_fro"], 1.5e-3) for r in rs]
    mratio = [r["cuda_max"] / max(r["eager_max"], 4e-3) for r in rs]
    w = max(rs, key=lambda r: r["cuda_fro"])
    ratio_sorted = sorted(ratio)
    print(f"| {k} | {len(rs)} | {w['cuda_fro']:.2e} | {w['eager_fro']:.2e} | {max(ratio):.2f} | {ratio_sorted[len(ratio) // 2]:.2f} | {max(mratio):.2f} |")
for r in recs:
    if r.get("what", "").startswith("greedy"):
        print(f"\nGreedy decoding, 32 tokens: CUDA == eager-bf16 oracle: {r['cuda'] == r['eager_bf16']}; CUDA == fp32 oracle: "
              f"{r['cuda'] == r['fp32']}; distinct tokens {r['distinct']}; min top-1/top-2 logit margin fp32 {r['min_margin_fp32']:.2f}.")
