"""B200-native mirror of the reference's `cambrian.model` package (the drop-in boundary)."""
from .language_model.cambrian_llama import CambrianConfig, CambrianLlamaForCausalLM, CambrianLlamaModel  # noqa: F401
