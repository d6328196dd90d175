"""State-dict specification (key -> shape) of the reference's D2SwinTransformer
(projects/IDOL/idol/backbone/swin.py:684-768; yaml projects/IDOL/configs/ytvis19_swinL.yaml:7-17 for Swin-L)."""

SWIN_L = dict(embed_dim=192, depths=(2, 2, 18, 2), num_heads=(6, 12, 24, 48), window_size=12)
SWIN_T = dict(embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), window_size=7)


def swin_state_spec(prefix, embed_dim, depths, num_heads, window_size, patch_size=4, mlp_ratio=4, with_index=True):
    spec = {prefix + "patch_embed.proj.weight": (embed_dim, 3, patch_size, patch_size),
            prefix + "patch_embed.proj.bias": (embed_dim,),
            prefix + "patch_embed.norm.weight": (embed_dim,), prefix + "patch_embed.norm.bias": (embed_dim,)}
    for i, (depth, heads) in enumerate(zip(depths, num_heads)):
        C = embed_dim * 2 ** i
        for j in range(depth):
            p = f"{prefix}layers.{i}.blocks.{j}."
            for n in ("norm1", "norm2"):
                spec[p + n + ".weight"] = (C,)
                spec[p + n + ".bias"] = (C,)
            spec[p + "attn.relative_position_bias_table"] = ((2 * window_size - 1) ** 2, heads)
            if with_index:
                spec[p + "attn.relative_position_index"] = ("index", window_size)
            spec[p + "attn.qkv.weight"] = (3 * C, C)
            spec[p + "attn.qkv.bias"] = (3 * C,)
            spec[p + "attn.proj.weight"] = (C, C)
            spec[p + "attn.proj.bias"] = (C,)
            spec[p + "mlp.fc1.weight"] = (mlp_ratio * C, C)
            spec[p + "mlp.fc1.bias"] = (mlp_ratio * C,)
            spec[p + "mlp.fc2.weight"] = (C, mlp_ratio * C)
            spec[p + "mlp.fc2.bias"] = (C,)
        if i < len(depths) - 1:
            p = f"{prefix}layers.{i}.downsample."
            spec[p + "reduction.weight"] = (2 * C, 4 * C)
            spec[p + "norm.weight"] = (4 * C,)
            spec[p + "norm.bias"] = (4 * C,)
        spec[f"{prefix}norm{i}.weight"] = (C,)
        spec[f"{prefix}norm{i}.bias"] = (C,)
    return spec
