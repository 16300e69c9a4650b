"""Host-side mirror of the reference's nn.Module interface for the hot path (same class names, constructor
arguments, forward signatures, parameter/buffer names and shapes — strict state_dict compatibility)."""
from .attention import Auto_Attn, ExampleGuidedAttention  # noqa: F401
