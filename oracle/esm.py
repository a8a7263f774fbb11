"""fp32 CPU restatement of the forward pass the reference triggers for ESM-2 checkpoints.

distllm's Esm2Encoder.encode (distllm/embed/encoders/esm2.py:109-134) calls
``EsmForMaskedLM(**batch, output_hidden_states=True)`` and returns ``hidden_states[-1]``, which is
the encoder output after ``emb_layer_norm_after`` (the LM head is irrelevant).  Restated from
transformers 5.5.0, transformers/models/esm/modeling_esm.py:

    embeddings   :189-234  word embedding, token dropout (<mask> rows zeroed, rescale by
                           (1 - 0.12) / (1 - observed mask ratio)), * attention_mask
    rotary       :43-54, :81-123   x cos + rotate_half(x) sin on q and k, angle = pos * 10000^(-2i/d)
    attention    :318-362  q scaled by d^-0.5 BEFORE the rotation, softmax(q k^T + padding mask) v
    blocks       :386-404, :446-483  pre-LayerNorm: x += attn(LN(x)); x += ffn(LN(x)), erf GELU (:57-61)
    final norm   :511-512  emb_layer_norm_after

Plain torch ops on CPU in fp32; the state dict uses HF parameter names.  TEST INFRASTRUCTURE ONLY.
"""

from __future__ import annotations

import math
from typing import Mapping

import torch
import torch.nn.functional as F  # noqa: N812


def _sd(state_dict: Mapping[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    return {
        (k[4:] if k.startswith('esm.') else k): v.detach().to('cpu', torch.float32)
        for k, v in state_dict.items()
    }


def _rotate(x: torch.Tensor) -> torch.Tensor:
    """x: [B, heads, S, d] -> rotary-embedded x."""
    d = x.shape[-1]
    s = x.shape[-2]
    inv_freq = 1.0 / (10000 ** (torch.arange(0, d, 2, dtype=torch.int64).float() / d))
    freqs = torch.outer(torch.arange(s).float(), inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos()[None, None], emb.sin()[None, None]
    x1, x2 = x.chunk(2, dim=-1)
    return x * cos + torch.cat((-x2, x1), dim=-1) * sin


@torch.no_grad()
def esm_forward(
    state_dict: Mapping[str, torch.Tensor],
    hf_config,
    input_ids: torch.Tensor,
    attention_mask: torch.Tensor,
    return_all: bool = False,
):
    """Last hidden state ``[B,S,H]`` fp32 (== ``EsmForMaskedLM(...).hidden_states[-1]``).

    ``return_all``: list over l = 1..L of ``emb_layer_norm_after(residual stream after l layers)`` -- what a
    model truncated to l layers would return (used by the per-layer drift report)."""
    sd = _sd(state_dict)
    eps = hf_config.layer_norm_eps
    heads = hf_config.num_attention_heads
    b, s = input_ids.shape
    h = hf_config.hidden_size
    d = h // heads

    x = sd['embeddings.word_embeddings.weight'][input_ids]
    if getattr(hf_config, 'token_dropout', False):
        is_mask = input_ids == hf_config.mask_token_id
        x = x.masked_fill(is_mask.unsqueeze(-1), 0.0)
        src_lengths = attention_mask.sum(-1)
        observed = is_mask.sum(-1).float() / src_lengths
        x = x * (1 - 0.15 * 0.8) / (1 - observed)[:, None, None]
    x = x * attention_mask.unsqueeze(-1).to(x.dtype)

    key_bias = torch.zeros(b, 1, 1, s)
    key_bias.masked_fill_(attention_mask.view(b, 1, 1, s) == 0, torch.finfo(torch.float32).min)

    states = []
    for layer in range(hf_config.num_hidden_layers):
        p = f'encoder.layer.{layer}.'

        def lin(t: torch.Tensor, name: str) -> torch.Tensor:
            return F.linear(t, sd[p + name + '.weight'], sd[p + name + '.bias'])

        def split(t: torch.Tensor) -> torch.Tensor:
            return t.view(b, s, heads, d).transpose(1, 2)

        y = F.layer_norm(x, (h,), sd[p + 'attention.LayerNorm.weight'], sd[p + 'attention.LayerNorm.bias'], eps)
        q = split(lin(y, 'attention.self.query')) * d ** -0.5
        k = split(lin(y, 'attention.self.key'))
        v = split(lin(y, 'attention.self.value'))
        q, k = _rotate(q), _rotate(k)
        scores = q @ k.transpose(-1, -2) + key_bias
        ctx = (torch.softmax(scores, dim=-1) @ v).transpose(1, 2).reshape(b, s, h)
        x = x + lin(ctx, 'attention.output.dense')
        y = F.layer_norm(x, (h,), sd[p + 'LayerNorm.weight'], sd[p + 'LayerNorm.bias'], eps)
        inter = lin(y, 'intermediate.dense')
        inter = inter * 0.5 * (1.0 + torch.erf(inter / math.sqrt(2.0)))
        x = x + lin(inter, 'output.dense')
        if return_all:
            states.append(F.layer_norm(x, (h,), sd['encoder.emb_layer_norm_after.weight'],
                                       sd['encoder.emb_layer_norm_after.bias'], eps))
    if return_all:
        return states
    return F.layer_norm(x, (h,), sd['encoder.emb_layer_norm_after.weight'],
                        sd['encoder.emb_layer_norm_after.bias'], eps)
