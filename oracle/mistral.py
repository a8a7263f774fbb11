"""fp32 CPU restatement of the forward pass the reference triggers for Mistral-family checkpoints.

distllm's AutoEncoder.encode (distllm/embed/encoders/auto.py:119-138) calls
``AutoModel(**batch, output_hidden_states=True)`` and returns ``hidden_states[-1]``; for a
``MistralModel`` that is the output of the final RMSNorm.  Restated from transformers 5.5.0,
transformers/models/mistral/modeling_mistral.py:

    MLP          :35-48    down(silu(gate(x)) * up(x)), no biases
    rotary       :51-80, :262-326  cos/sin of pos * theta^(-2i/d), halves convention (rotate_half)
    attention    :122-180  q/k/v projections without bias, grouped-query (kv heads repeated), scores
                           scaled by d^-0.5, causal + sliding-window + key-padding mask, o_proj
    RMSNorm      :181-200  x * rsqrt(mean(x^2) + eps) * weight, statistics in fp32
    blocks       :202-242  pre-norm: x += attn(norm(x)); x += mlp(norm(x))
    model        :328-400  embed_tokens, position_ids = arange(S) for every row (padding does not
                           shift positions), final norm
    mask         transformers/masking_utils.py (sliding_window_causal): key j is visible to query i
                 iff j <= i, i - j < sliding_window, and attention_mask[b, j] != 0

Query rows that see no key at all (left padding) come out of torch SDPA as zeros; such rows are
never read by later valid tokens nor by the poolers, and this restatement reproduces the zeros.
Plain torch ops on CPU in fp32; the state dict uses HF parameter names.  TEST INFRASTRUCTURE ONLY.
"""

from __future__ import annotations

from typing import Mapping

import torch
import torch.nn.functional as F  # noqa: N812


class _LazyF32:
    """State-dict view that strips the ``model.`` prefix and converts a tensor to CPU fp32 only when it is
    read: a 7B-parameter checkpoint (28 GB as fp32) is then walked one projection at a time, e.g. straight
    from bf16 device tensors."""

    def __init__(self, state_dict: Mapping[str, torch.Tensor]) -> None:
        self._sd = {(k[6:] if k.startswith('model.') else k): v for k, v in state_dict.items()}

    def __getitem__(self, key: str) -> torch.Tensor:
        return self._sd[key].detach().to('cpu', torch.float32)


def _sd(state_dict: Mapping[str, torch.Tensor]) -> _LazyF32:
    return _LazyF32(state_dict)


def rope_theta_of(hf_config) -> float:
    params = getattr(hf_config, 'rope_parameters', None)
    if params and 'rope_theta' in params:
        return float(params['rope_theta'])
    return float(getattr(hf_config, 'rope_theta', 10000.0))


def _rms(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    return w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))


def _rotate(x: torch.Tensor, theta: float) -> torch.Tensor:
    """x: [B, heads, S, d] -> rotary-embedded x (positions 0..S-1)."""
    d, s = x.shape[-1], x.shape[-2]
    inv_freq = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.int64).float() / d))
    freqs = torch.outer(torch.arange(s).float(), inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos()[None, None], emb.sin()[None, None]
    x1, x2 = x.chunk(2, dim=-1)
    return x * cos + torch.cat((-x2, x1), dim=-1) * sin


def visibility(attention_mask: torch.Tensor, sliding_window: int | None) -> torch.Tensor:
    """bool [B,1,S,S]: may query i read key j."""
    b, s = attention_mask.shape
    i = torch.arange(s)[:, None]
    j = torch.arange(s)[None, :]
    vis = j <= i
    if sliding_window is not None:
        vis = vis & (i - j < sliding_window)
    return vis[None, None] & (attention_mask != 0)[:, None, None, :]


@torch.no_grad()
def mistral_forward(
    state_dict: Mapping[str, torch.Tensor],
    hf_config,
    input_ids: torch.Tensor,
    attention_mask: torch.Tensor,
    return_all: bool = False,
):
    """Last hidden state ``[B,S,H]`` fp32 (== ``MistralModel(...).hidden_states[-1]``).

    ``return_all``: list over l = 1..L of ``final_norm(residual stream after l layers)`` -- what a model
    truncated to l layers would return (the per-layer drift report compares against these)."""
    sd = _sd(state_dict)
    eps = hf_config.rms_norm_eps
    heads, kv_heads = hf_config.num_attention_heads, hf_config.num_key_value_heads
    b, s = input_ids.shape
    h = hf_config.hidden_size
    d = getattr(hf_config, 'head_dim', None) or h // heads
    theta = rope_theta_of(hf_config)
    vis = visibility(attention_mask, getattr(hf_config, 'sliding_window', None))
    dead = ~vis.any(-1, keepdim=True)  # [B,1,S,1] query rows without a visible key

    x = sd['embed_tokens.weight'][input_ids]
    states = []
    for layer in range(hf_config.num_hidden_layers):
        p = f'layers.{layer}.'
        y = _rms(x, sd[p + 'input_layernorm.weight'], eps)
        q = F.linear(y, sd[p + 'self_attn.q_proj.weight']).view(b, s, heads, d).transpose(1, 2)
        k = F.linear(y, sd[p + 'self_attn.k_proj.weight']).view(b, s, kv_heads, d).transpose(1, 2)
        v = F.linear(y, sd[p + 'self_attn.v_proj.weight']).view(b, s, kv_heads, d).transpose(1, 2)
        q, k = _rotate(q, theta), _rotate(k, theta)
        k = k.repeat_interleave(heads // kv_heads, dim=1)
        v = v.repeat_interleave(heads // kv_heads, dim=1)
        scores = (q @ k.transpose(-1, -2)) * d ** -0.5
        scores = scores.masked_fill(~vis, float('-inf'))
        prob = torch.softmax(scores, dim=-1).masked_fill(dead, 0.0)
        ctx = (prob @ v).transpose(1, 2).reshape(b, s, heads * d)
        x = x + F.linear(ctx, sd[p + 'self_attn.o_proj.weight'])
        y = _rms(x, sd[p + 'post_attention_layernorm.weight'], eps)
        gate = F.linear(y, sd[p + 'mlp.gate_proj.weight'])
        up = F.linear(y, sd[p + 'mlp.up_proj.weight'])
        x = x + F.linear(F.silu(gate) * up, sd[p + 'mlp.down_proj.weight'])
        if return_all:
            states.append(_rms(x, sd['norm.weight'], eps))
    return states if return_all else _rms(x, sd['norm.weight'], eps)
