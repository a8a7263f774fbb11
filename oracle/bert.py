"""fp32 CPU restatement of the forward pass the reference triggers for BERT checkpoints.

distllm's AutoEncoder.encode (distllm/embed/encoders/auto.py:119-138) calls
``AutoModel(**batch, output_hidden_states=True)`` and returns ``hidden_states[-1]``; for a BERT
checkpoint that is HF ``BertModel.forward`` (transformers 5.5.0,
transformers/models/bert/modeling_bert.py):

    embeddings   :72-112   (word + token_type) + position -> LayerNorm
    self-attn    :168-207  q/k/v Linear, softmax(q k^T / sqrt(d) + padding mask) v
    self-output  :294-298  Linear + residual -> LayerNorm
    intermediate :339-342  Linear -> erf GELU
    output       :352-356  Linear + residual -> LayerNorm
    padding mask :692-716  additive, most-negative-finite on padded keys

Plain torch ops on CPU in fp32; the state dict uses HF parameter names.  TEST INFRASTRUCTURE ONLY.
"""

from __future__ import annotations

import math
from typing import Mapping

import torch
import torch.nn.functional as F  # noqa: N812


def _sd(state_dict: Mapping[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    return {
        (k[5:] if k.startswith('bert.') else k): v.detach().to('cpu', torch.float32)
        for k, v in state_dict.items()
    }


@torch.no_grad()
def bert_forward(
    state_dict: Mapping[str, torch.Tensor],
    hf_config,
    input_ids: torch.Tensor,
    attention_mask: torch.Tensor,
    token_type_ids: torch.Tensor | None = None,
    return_all: bool = False,
):
    """Last hidden state ``[B,S,H]`` fp32 (== ``outputs.hidden_states[-1]``)."""
    sd = _sd(state_dict)
    eps = hf_config.layer_norm_eps
    heads = hf_config.num_attention_heads
    b, s = input_ids.shape
    h = hf_config.hidden_size
    d = h // heads
    if token_type_ids is None:
        token_type_ids = torch.zeros_like(input_ids)

    x = sd['embeddings.word_embeddings.weight'][input_ids]
    x = x + sd['embeddings.token_type_embeddings.weight'][token_type_ids]
    x = x + sd['embeddings.position_embeddings.weight'][torch.arange(s)][None]
    x = F.layer_norm(x, (h,), sd['embeddings.LayerNorm.weight'], sd['embeddings.LayerNorm.bias'], eps)

    key_bias = torch.zeros(b, 1, 1, s)
    key_bias.masked_fill_(attention_mask.view(b, 1, 1, s) == 0, torch.finfo(torch.float32).min)

    states = [x]
    for layer in range(hf_config.num_hidden_layers):
        p = f'encoder.layer.{layer}.'

        def lin(t: torch.Tensor, name: str) -> torch.Tensor:
            return F.linear(t, sd[p + name + '.weight'], sd[p + name + '.bias'])

        def split(t: torch.Tensor) -> torch.Tensor:
            return t.view(b, s, heads, d).transpose(1, 2)

        q, k, v = (split(lin(x, f'attention.self.{n}')) for n in ('query', 'key', 'value'))
        scores = q @ k.transpose(-1, -2) / math.sqrt(d) + key_bias
        ctx = (torch.softmax(scores, dim=-1) @ v).transpose(1, 2).reshape(b, s, h)
        x = F.layer_norm(lin(ctx, 'attention.output.dense') + x, (h,),
                         sd[p + 'attention.output.LayerNorm.weight'],
                         sd[p + 'attention.output.LayerNorm.bias'], eps)
        inter = F.gelu(lin(x, 'intermediate.dense'))
        x = F.layer_norm(lin(inter, 'output.dense') + x, (h,), sd[p + 'output.LayerNorm.weight'],
                         sd[p + 'output.LayerNorm.bias'], eps)
        states.append(x)
    return states if return_all else x
