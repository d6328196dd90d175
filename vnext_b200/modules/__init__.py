from .ms_deform_attn import MSDeformAttn, MSDeformAttnSeqFormer  # noqa: F401
