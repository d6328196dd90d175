"""IDOL (online VIS) on the B200 hot path: per-frame Deformable-DETR + CondInst mask head + reID embedding head."""
