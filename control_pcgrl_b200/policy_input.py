"""Policy-side input for a network that lives on the env's GPU.

The reference's policies take the float64 one-hot crop and start with
`input_dict["obs"].permute(0, 3, 1, 2).float()` (control_pcgrl/rl/models.py:60-66: RLlib hands channel-last
observations, torch convolutions want NCHW).  With the env on the same device the observation never has to exist in
that form in HBM: `BatchedPcgrlEnv.observe(onehot=False)` writes one uint8 tile code per pixel (0 = outside the map,
tile t -> t + 1, which is exactly the one-hot channel index of Cropped + OneHotEncoding, wrappers.py:407-437,
232-257), 1 byte per pixel instead of 8 * (n_tiles + 1), and the expansion to the network's input layout happens in
the same pass that feeds the first convolution.
"""
from __future__ import annotations

import torch


def conv_input_from_codes(codes: torch.Tensor, n_channels: int, dtype=torch.float32, channels_last: bool = True):
    """codes: [N, *dims, 1] uint8 from `observe(onehot=False)`; n_channels = n_tiles + 1 for a cropped observation
    (n_tiles for the wide / cellular ones).  Returns the [N, n_channels, *dims] tensor the reference's first
    convolution sees (`obs.permute(0, 3, 1, 2).float()` of the one-hot observation), in `dtype`; for 2D inputs in
    torch's channels_last memory format by default, which is the layout the scatter below writes sequentially."""
    if codes.dtype != torch.uint8 or codes.shape[-1] != 1:
        raise ValueError("codes must be the uint8 [N, *dims, 1] tensor of observe(onehot=False)")
    idx = codes.to(torch.int64)                                   # [N, *dims, 1]
    onehot = torch.zeros((*codes.shape[:-1], n_channels), dtype=dtype, device=codes.device)
    onehot.scatter_(-1, idx, 1)                                   # channel-last one-hot: the reference's observation
    nd = codes.dim() - 2
    out = onehot.permute(0, nd + 1, *range(1, nd + 1))            # NCHW view of NHWC storage == channels_last
    if not channels_last or nd != 2:
        out = out.contiguous()
    return out
