"""Seeded synthetic inputs shared by the golden generator, the parity tests and the bench
(SURVEY.md §8d 'Synthetic inputs')."""
import torch


def make_inputs(cfg, B, lat, L, seed, n_controlnets=2):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, cfg.in_channels, lat, lat, generator=g)
    ctx = torch.randn(B, L, cfg.context_dim, generator=g)
    hints = []
    # SAM id map: integer valued 0..255, channel 2 zero, NOT normalised (editany_lora.py:444-446,776)
    h0 = torch.randint(0, 256, (1, cfg.hint_channels, 8 * lat, 8 * lat), generator=g).float()
    h0[:, 2] = 0
    hints.append(h0.repeat(B, 1, 1, 1))
    if n_controlnets > 1:
        # inpaint condition: image in [0,1], masked square := -1 (editany_lora.py:332-340)
        h1 = torch.rand(1, cfg.hint_channels, 8 * lat, 8 * lat, generator=g)
        q = 2 * lat
        h1[:, :, q:8 * lat - q, q:8 * lat - q] = -1.0
        hints.append(h1.repeat(B, 1, 1, 1))
    return x, ctx, hints[:n_controlnets]
