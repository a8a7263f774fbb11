"""NF4 weight quantisation as the reference's default ``quantization=True`` applies it.

distllm/embed/encoders/auto.py:44-56 loads the checkpoint with
``BitsAndBytesConfig(load_in_4bit=True, bnb_4bit_quant_type='nf4', bnb_4bit_use_double_quant=True,
bnb_4bit_compute_dtype=torch.bfloat16)``: every ``nn.Linear`` weight of the model is stored as 4-bit NormalFloat
codes and DEQUANTISED to the compute dtype in front of each matmul.  What reaches the GEMMs is therefore
``dequant(quant(W))``; this module computes exactly that tensor once, at load time, and hands it to the native
16-bit GEMMs (the 4-bit storage itself -- a memory saving, not an arithmetic one -- is not reproduced).

bitsandbytes (pin >=0.42.0, pyproject.toml) is absent from this image and cannot run on CPU, so this restates
its published algorithm -- PARITY UNPINNED for this branch:

  quantize_4bit(W, blocksize=64, quant_type='nf4')   flatten; per block of 64 values: absmax = max|w|; each
      w / absmax is replaced by the nearest of the 16 NF4 code values (quantiles of N(0,1) normalised to [-1, 1])
  compress_statistics (double quantisation)           offset = mean(absmax); absmax - offset is quantised
      blockwise (blocks of 256) to the 256-entry signed "dynamic" 8-bit code: per block absmax2 = max|x|, each
      x / absmax2 -> nearest code value
  dequantize_4bit                                      absmax' = code8[q] * absmax2 + offset;  w' = nf4[q4] * absmax'
"""

from __future__ import annotations

import torch

NF4_CODE = (
    -1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
    -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
    0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0,
)


def dynamic_map_8bit() -> torch.Tensor:
    """bitsandbytes.functional.create_dynamic_map(signed=True, max_exponent_bits=7, total_bits=8): 2^i values per
    decade 10^(i-6) (i = 0..6, midpoints of a linear grid over [0.1, 1]) of either sign, plus 0 and 1."""
    data: list[float] = []
    for i in range(7):
        boundaries = torch.linspace(0.1, 1, 2 ** i + 1)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        scale = 10.0 ** (-6 + i)
        data += (scale * means).tolist()
        data += (-scale * means).tolist()
    data += [0.0, 1.0]
    assert len(data) == 256
    return torch.tensor(sorted(data), dtype=torch.float32)


def _nearest(values: torch.Tensor, code: torch.Tensor) -> torch.Tensor:
    """Index of the nearest code value (code sorted ascending) for every element."""
    mid = (code[:-1] + code[1:]) / 2.0
    return torch.bucketize(values, mid.to(values.device))


@torch.no_grad()
def nf4_roundtrip(weight: torch.Tensor, blocksize: int = 64, double_quant: bool = True) -> torch.Tensor:
    """``dequantize_4bit(quantize_4bit(weight))`` as fp32, same shape and device as ``weight``."""
    w = weight.detach().to(torch.float32)
    flat = w.flatten()
    n = flat.numel()
    pad = (-n) % blocksize
    if pad:
        flat = torch.cat([flat, flat.new_zeros(pad)])
    blocks = flat.view(-1, blocksize)
    absmax = blocks.abs().amax(dim=1)
    code = torch.tensor(NF4_CODE, dtype=torch.float32, device=w.device)
    safe = torch.where(absmax > 0, absmax, torch.ones_like(absmax))
    q4 = _nearest(blocks / safe[:, None], code)
    if double_quant:
        code8 = dynamic_map_8bit().to(w.device)
        offset = absmax.mean()
        centred = absmax - offset
        pad2 = (-centred.numel()) % 256
        c = torch.cat([centred, centred.new_zeros(pad2)]) if pad2 else centred
        c = c.view(-1, 256)
        absmax2 = c.abs().amax(dim=1)
        safe2 = torch.where(absmax2 > 0, absmax2, torch.ones_like(absmax2))
        q8 = _nearest(c / safe2[:, None], code8)
        absmax = (code8[q8] * absmax2[:, None]).flatten()[: absmax.numel()] + offset
    out = (code[q4] * absmax[:, None]).flatten()[:n]
    return out.view_as(w)


_LINEAR_SUFFIXES = (
    # BERT / ESM-2
    'attention.self.query.weight', 'attention.self.key.weight', 'attention.self.value.weight',
    'attention.output.dense.weight', 'intermediate.dense.weight', 'output.dense.weight',
    # Mistral
    'self_attn.q_proj.weight', 'self_attn.k_proj.weight', 'self_attn.v_proj.weight', 'self_attn.o_proj.weight',
    'mlp.gate_proj.weight', 'mlp.up_proj.weight', 'mlp.down_proj.weight',
    # ModernBERT
    'attn.Wqkv.weight', 'attn.Wo.weight', 'mlp.Wi.weight', 'mlp.Wo.weight',
)


def quantize_state_dict_nf4(state_dict: dict[str, torch.Tensor], device: torch.device | str | None = None) -> dict:
    """Copy of ``state_dict`` whose transformer-block ``nn.Linear`` weights went through NF4 (embeddings, norms
    and biases are not quantised by bitsandbytes either).  The round trip runs on ``device`` when given."""
    out = {}
    for name, t in state_dict.items():
        if name.endswith(_LINEAR_SUFFIXES) and t.dim() == 2:
            src = t.to(device) if device is not None else t
            out[name] = nf4_roundtrip(src)
        else:
            out[name] = t
    return out
