"""fp32 CPU restatement of the forward pass the reference triggers for ModernBERT checkpoints.

distllm's AutoEncoder.encode (distllm/embed/encoders/auto.py:119-138) calls
``AutoModel(**batch, output_hidden_states=True)`` and returns ``hidden_states[-1]``; for a ``modernbert``
checkpoint (examples/embed/workstation/modernbert_semchunk.yaml:16-17) that is the output of ``final_norm`` of
HF ``ModernBertModel``.  Restated from transformers 5.5.0, transformers/models/modernbert/modeling_modernbert.py:

    embeddings   :52-71    tok_embeddings -> LayerNorm (no position table: rotary)
    MLP          :74-91    Wi -> chunk(2): act(input) * gate -> Wo          (GeGLU, erf GELU)
    rotary       :94-175, :201-229  per layer TYPE: theta 160000 on full-attention layers, 10000 on sliding ones;
                           halves convention (rotate_half), positions 0..S-1 for every row
    attention    :232-310  fused Wqkv (q | k | v), heads of hidden/heads, scores * d^-0.5, bidirectional
    block        :313-343  pre-norm: x += attn(attn_norm(x)); x += mlp(mlp_norm(x)); layer 0 has NO attn_norm
    model        :424-490  final_norm; masks: key padding everywhere, and on sliding layers additionally
                           |i - j| <= config.sliding_window (= local_attention // 2; masking_utils.py
                           sliding_window_bidirectional_overlay: inclusive distance)

Biases (norm_bias / attention_bias / mlp_bias) are honoured when the state dict has them.  Plain torch ops on
CPU in fp32; HF parameter names.  TEST INFRASTRUCTURE ONLY.
"""

from __future__ import annotations

import math
from typing import Mapping

import torch
import torch.nn.functional as F  # noqa: N812


def _sd(state_dict: Mapping[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    return {
        (k[6:] if k.startswith('model.') else k): v.detach().to('cpu', torch.float32)
        for k, v in state_dict.items()
    }


def layer_is_global(hf_config, layer: int) -> bool:
    return hf_config.layer_types[layer] == 'full_attention'


def rope_thetas(hf_config) -> tuple[float, float]:
    """(theta of full-attention layers, theta of sliding-attention layers)."""
    params = hf_config.rope_parameters
    return float(params['full_attention']['rope_theta']), float(params['sliding_attention']['rope_theta'])


def _rotate(x: torch.Tensor, theta: float) -> torch.Tensor:
    d, s = x.shape[-1], x.shape[-2]
    inv_freq = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.int64).float() / d))
    freqs = torch.outer(torch.arange(s).float(), inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos()[None, None], emb.sin()[None, None]
    x1, x2 = x.chunk(2, dim=-1)
    return x * cos + torch.cat((-x2, x1), dim=-1) * sin


@torch.no_grad()
def modernbert_forward(
    state_dict: Mapping[str, torch.Tensor],
    hf_config,
    input_ids: torch.Tensor,
    attention_mask: torch.Tensor,
    return_all: bool = False,
):
    """Last hidden state ``[B,S,H]`` fp32 (== ``ModernBertModel(...).last_hidden_state``).

    ``return_all``: list over l = 1..L of ``final_norm(residual stream after l layers)``."""
    sd = _sd(state_dict)
    eps = hf_config.norm_eps
    heads = hf_config.num_attention_heads
    b, s = input_ids.shape
    h = hf_config.hidden_size
    d = h // heads
    theta_global, theta_local = rope_thetas(hf_config)
    window = int(hf_config.sliding_window)

    def ln(x: torch.Tensor, name: str) -> torch.Tensor:
        return F.layer_norm(x, (h,), sd[name + '.weight'], sd.get(name + '.bias'), eps)

    def lin(x: torch.Tensor, name: str) -> torch.Tensor:
        return F.linear(x, sd[name + '.weight'], sd.get(name + '.bias'))

    x = ln(sd['embeddings.tok_embeddings.weight'][input_ids], 'embeddings.norm')
    key_ok = (attention_mask != 0).view(b, 1, 1, s)
    i = torch.arange(s)
    band = ((i[:, None] - i[None, :]).abs() <= window)[None, None]
    neg = torch.finfo(torch.float32).min
    bias_global = torch.zeros(b, 1, 1, s).masked_fill(~key_ok, neg)
    bias_local = torch.zeros(b, 1, s, s).masked_fill(~(key_ok & band), neg)

    states = []
    for layer in range(hf_config.num_hidden_layers):
        p = f'layers.{layer}.'
        is_global = layer_is_global(hf_config, layer)
        y = x if layer == 0 else ln(x, p + 'attn_norm')
        qkv = lin(y, p + 'attn.Wqkv').view(b, s, 3, heads, d)
        q, k, v = (qkv[:, :, j].transpose(1, 2) for j in range(3))
        theta = theta_global if is_global else theta_local
        q, k = _rotate(q, theta), _rotate(k, theta)
        scores = q @ k.transpose(-1, -2) * d ** -0.5 + (bias_global if is_global else bias_local)
        ctx = (torch.softmax(scores, dim=-1) @ v).transpose(1, 2).reshape(b, s, h)
        x = x + lin(ctx, p + 'attn.Wo')
        y = ln(x, p + 'mlp_norm')
        inp, gate = lin(y, p + 'mlp.Wi').chunk(2, dim=-1)
        act = inp * 0.5 * (1.0 + torch.erf(inp / math.sqrt(2.0)))
        x = x + lin(act * gate, p + 'mlp.Wo')
        if return_all:
            states.append(ln(x, 'final_norm'))
    return states if return_all else ln(x, 'final_norm')
