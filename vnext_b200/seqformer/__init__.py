"""SeqFormer (offline VIS) on the B200 hot path: state-dict layout, inference engine, META_ARCH wrapper."""
