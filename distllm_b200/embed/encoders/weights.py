"""Weight layout handed to ``b2e_encoder_create`` (the order IS the ABI, see include/b2e.h).

BERT (``B2E_ARCH_BERT``), 5 + 12*L device tensors:

    0 word_embeddings [V,H] f32      1 position_embeddings [P,H] f32   2 token_type_embeddings [T,H] f32
    3 embeddings.LayerNorm.weight    4 embeddings.LayerNorm.bias       (f32 [H])
    per layer l, base = 5 + 12*l:
      +0 Wqkv [3H,H] 16-bit (rows: query | key | value)   +1 bqkv [3H] f32
      +2 Wo   [H,H]  16-bit                               +3 bo   [H]  f32
      +4 attention.output.LayerNorm.weight  +5 .bias    (f32)
      +6 W1   [I,H]  16-bit (intermediate.dense)          +7 b1   [I]  f32
      +8 W2   [H,I]  16-bit (output.dense)                +9 b2   [H]  f32
      +10 output.LayerNorm.weight           +11 .bias   (f32)

Names on the right are HF ``BertModel`` state-dict keys (transformers/models/bert/modeling_bert.py).
Matrices keep nn.Linear's [out_features, in_features] layout, which is the K-major B operand the
tcgen05 GEMM wants, so no transposes are needed.
"""

from __future__ import annotations

from typing import Mapping

import torch

from distllm_b200 import _native


HALF_MAX = 65504.0


def to_storage(t: torch.Tensor, device: torch.device, dtype: torch.dtype) -> torch.Tensor:
    """Checkpoint matrix (fp32 / bf16 / fp16) -> contiguous device tensor of the library build's 16-bit storage
    type.  float16 saturates at +-65504 (weights never get near it; the clamp only keeps a broken checkpoint
    from turning into inf)."""
    x = t.detach().to(device=device, dtype=torch.float32)
    if dtype == torch.float16:
        x = x.clamp(-HALF_MAX, HALF_MAX)
    return x.to(dtype).contiguous()


def bert_desc(hf_config) -> _native.ModelDesc:
    """Translate a HF ``BertConfig`` into the C ``B2EModelDesc``; reject what is not built."""
    if getattr(hf_config, 'position_embedding_type', 'absolute') not in (None, 'absolute'):
        raise NotImplementedError('only absolute position embeddings are supported')
    act = getattr(hf_config, 'hidden_act', 'gelu')
    if act != 'gelu':
        raise NotImplementedError(f"hidden_act={act!r}: only erf-GELU ('gelu') is built")
    heads = hf_config.num_attention_heads
    return _native.ModelDesc(
        arch=_native.ARCH_BERT,
        num_layers=hf_config.num_hidden_layers,
        hidden=hf_config.hidden_size,
        heads=heads,
        kv_heads=heads,
        head_dim=hf_config.hidden_size // heads,
        intermediate=hf_config.intermediate_size,
        vocab=hf_config.vocab_size,
        max_pos=hf_config.max_position_embeddings,
        type_vocab=hf_config.type_vocab_size,
        eps=float(hf_config.layer_norm_eps),
        rope_theta=0.0,
        sliding_window=0,
        reserved=0,
    )


def bert_weight_list(
    state_dict: Mapping[str, torch.Tensor],
    num_layers: int,
    device: torch.device,
    dtype: torch.dtype = torch.float16,
) -> list[torch.Tensor]:
    """HF BertModel state dict -> contiguous device tensors in ABI order."""
    sd = {k[5:] if k.startswith('bert.') else k: v for k, v in state_dict.items()}

    def f32(key: str) -> torch.Tensor:
        return sd[key].detach().to(device=device, dtype=torch.float32).contiguous()

    def b16(t: torch.Tensor) -> torch.Tensor:
        return to_storage(t, device, dtype)

    out = [
        f32('embeddings.word_embeddings.weight'),
        f32('embeddings.position_embeddings.weight'),
        f32('embeddings.token_type_embeddings.weight'),
        f32('embeddings.LayerNorm.weight'),
        f32('embeddings.LayerNorm.bias'),
    ]
    for layer in range(num_layers):
        p = f'encoder.layer.{layer}.'
        qkv_w = torch.cat([sd[p + f'attention.self.{n}.weight'] for n in ('query', 'key', 'value')])
        qkv_b = torch.cat([sd[p + f'attention.self.{n}.bias'] for n in ('query', 'key', 'value')])
        out += [
            b16(qkv_w),
            qkv_b.detach().to(device=device, dtype=torch.float32).contiguous(),
            b16(sd[p + 'attention.output.dense.weight']),
            f32(p + 'attention.output.dense.bias'),
            f32(p + 'attention.output.LayerNorm.weight'),
            f32(p + 'attention.output.LayerNorm.bias'),
            b16(sd[p + 'intermediate.dense.weight']),
            f32(p + 'intermediate.dense.bias'),
            b16(sd[p + 'output.dense.weight']),
            f32(p + 'output.dense.bias'),
            f32(p + 'output.LayerNorm.weight'),
            f32(p + 'output.LayerNorm.bias'),
        ]
    return out


def random_bert_state_dict(hf_config, seed: int = 0, device: torch.device | str = 'cpu',
                           std: float | None = None) -> dict[str, torch.Tensor]:
    """Seeded random weights with HF BertModel names/shapes (normal(0, initializer_range),
    LayerNorm weight 1 / bias 0 -- HF's ``_init_weights``), generated directly on ``device``.

    Used for synthetic-weight benchmarking where no checkpoint can be downloaded.
    """
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    std = hf_config.initializer_range if std is None else std
    h, i = hf_config.hidden_size, hf_config.intermediate_size

    def normal(*shape: int) -> torch.Tensor:
        return torch.randn(*shape, generator=gen, device=device, dtype=torch.float32) * std

    sd = {
        'embeddings.word_embeddings.weight': normal(hf_config.vocab_size, h),
        'embeddings.position_embeddings.weight': normal(hf_config.max_position_embeddings, h),
        'embeddings.token_type_embeddings.weight': normal(hf_config.type_vocab_size, h),
        'embeddings.LayerNorm.weight': torch.ones(h, device=device),
        'embeddings.LayerNorm.bias': torch.zeros(h, device=device),
    }
    for layer in range(hf_config.num_hidden_layers):
        p = f'encoder.layer.{layer}.'
        for name, (o, k) in {
            'attention.self.query': (h, h),
            'attention.self.key': (h, h),
            'attention.self.value': (h, h),
            'attention.output.dense': (h, h),
            'intermediate.dense': (i, h),
            'output.dense': (h, i),
        }.items():
            sd[p + name + '.weight'] = normal(o, k)
            # HF zero-initialises biases; small non-zero values exercise the bias epilogues
            sd[p + name + '.bias'] = normal(o)
        for name in ('attention.output.LayerNorm', 'output.LayerNorm'):
            sd[p + name + '.weight'] = torch.ones(h, device=device)
            sd[p + name + '.bias'] = torch.zeros(h, device=device)
    return sd


# --------------------------------------------------------------------------- ESM-2
# ``B2E_ARCH_ESM2``, 3 + 12*L device tensors (HF EsmModel / EsmForMaskedLM names, ``esm.`` prefix
# stripped; transformers/models/esm/modeling_esm.py):
#
#     0 embeddings.word_embeddings [V,H] f32
#     1 encoder.emb_layer_norm_after.weight   2 .bias                         (f32 [H])
#     per layer l, base = 3 + 12*l (pre-LayerNorm blocks):
#       +0 attention.LayerNorm.weight  +1 .bias                 (LN before self-attention)
#       +2 Wqkv [3H,H] 16-bit (query | key | value)               +3 bqkv [3H] f32
#       +4 attention.output.dense.weight [H,H] 16-bit             +5 .bias
#       +6 LayerNorm.weight            +7 .bias                 (LN before the feed-forward)
#       +8 intermediate.dense.weight [I,H] 16-bit                 +9 .bias
#       +10 output.dense.weight [H,I] 16-bit                      +11 .bias
#
# ``B2EModelDesc.reserved`` carries ``mask_token_id + 1`` when ``token_dropout`` is on (0 = off).


def esm_desc(hf_config) -> _native.ModelDesc:
    """Translate a HF ``EsmConfig`` (ESM-2 family) into the C ``B2EModelDesc``."""
    if getattr(hf_config, 'position_embedding_type', 'absolute') != 'rotary':
        raise NotImplementedError('only rotary ESM-2 checkpoints are supported')
    if getattr(hf_config, 'emb_layer_norm_before', False):
        raise NotImplementedError('emb_layer_norm_before=True (ESM-1b style) is not built')
    heads = hf_config.num_attention_heads
    token_dropout = bool(getattr(hf_config, 'token_dropout', False))
    return _native.ModelDesc(
        arch=_native.ARCH_ESM2,
        num_layers=hf_config.num_hidden_layers,
        hidden=hf_config.hidden_size,
        heads=heads,
        kv_heads=heads,
        head_dim=hf_config.hidden_size // heads,
        intermediate=hf_config.intermediate_size,
        vocab=hf_config.vocab_size,
        max_pos=hf_config.max_position_embeddings,
        type_vocab=0,
        eps=float(hf_config.layer_norm_eps),
        rope_theta=10000.0,
        sliding_window=0,
        reserved=(int(hf_config.mask_token_id) + 1) if token_dropout else 0,
    )


def esm_weight_list(
    state_dict: Mapping[str, torch.Tensor],
    num_layers: int,
    device: torch.device,
    dtype: torch.dtype = torch.float16,
) -> list[torch.Tensor]:
    """HF EsmModel/EsmForMaskedLM state dict -> contiguous device tensors in ABI order."""
    sd = {k[4:] if k.startswith('esm.') else k: v for k, v in state_dict.items()}

    def f32(key: str) -> torch.Tensor:
        return sd[key].detach().to(device=device, dtype=torch.float32).contiguous()

    def b16(t: torch.Tensor) -> torch.Tensor:
        return to_storage(t, device, dtype)

    out = [
        f32('embeddings.word_embeddings.weight'),
        f32('encoder.emb_layer_norm_after.weight'),
        f32('encoder.emb_layer_norm_after.bias'),
    ]
    for layer in range(num_layers):
        p = f'encoder.layer.{layer}.'
        qkv_w = torch.cat([sd[p + f'attention.self.{n}.weight'] for n in ('query', 'key', 'value')])
        qkv_b = torch.cat([sd[p + f'attention.self.{n}.bias'] for n in ('query', 'key', 'value')])
        out += [
            f32(p + 'attention.LayerNorm.weight'),
            f32(p + 'attention.LayerNorm.bias'),
            b16(qkv_w),
            qkv_b.detach().to(device=device, dtype=torch.float32).contiguous(),
            b16(sd[p + 'attention.output.dense.weight']),
            f32(p + 'attention.output.dense.bias'),
            f32(p + 'LayerNorm.weight'),
            f32(p + 'LayerNorm.bias'),
            b16(sd[p + 'intermediate.dense.weight']),
            f32(p + 'intermediate.dense.bias'),
            b16(sd[p + 'output.dense.weight']),
            f32(p + 'output.dense.bias'),
        ]
    return out


def random_esm_state_dict(hf_config, seed: int = 0, device: torch.device | str = 'cpu',
                          std: float | None = None) -> dict[str, torch.Tensor]:
    """Seeded random ESM-2 weights with HF EsmModel names (no ``esm.`` prefix)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    std = getattr(hf_config, 'initializer_range', 0.02) if std is None else std
    h, i = hf_config.hidden_size, hf_config.intermediate_size

    def normal(*shape: int) -> torch.Tensor:
        return torch.randn(*shape, generator=gen, device=device, dtype=torch.float32) * std

    def ln(prefix: str, sd: dict) -> None:
        # non-trivial LayerNorm parameters so that both gamma and beta paths are exercised
        sd[prefix + '.weight'] = 1.0 + normal(h)
        sd[prefix + '.bias'] = normal(h)

    sd: dict[str, torch.Tensor] = {'embeddings.word_embeddings.weight': normal(hf_config.vocab_size, h)}
    ln('encoder.emb_layer_norm_after', sd)
    for layer in range(hf_config.num_hidden_layers):
        p = f'encoder.layer.{layer}.'
        for name, (o, k) in {
            'attention.self.query': (h, h),
            'attention.self.key': (h, h),
            'attention.self.value': (h, h),
            'attention.output.dense': (h, h),
            'intermediate.dense': (i, h),
            'output.dense': (h, i),
        }.items():
            sd[p + name + '.weight'] = normal(o, k)
            sd[p + name + '.bias'] = normal(o)
        ln(p + 'attention.LayerNorm', sd)
        ln(p + 'LayerNorm', sd)
    return sd


# ------------------------------------------------------------------------------ Mistral family
# Mistral (``B2E_ARCH_MISTRAL``), 2 + 6*L device tensors (no biases anywhere):
#
#     0 embed_tokens [V,H] f32          1 norm.weight [H] f32 (final RMSNorm)
#     per layer l, base = 2 + 6*l:
#       +0 input_layernorm.weight [H] f32
#       +1 Wqkv [(heads + 2*kv_heads)*d, H] 16-bit (rows: q_proj | k_proj | v_proj)
#       +2 Wo   [H, heads*d] 16-bit
#       +3 post_attention_layernorm.weight [H] f32
#       +4 Wgu  [2I, H] 16-bit: gate_proj and up_proj interleaved in blocks of 64 rows
#               (rows [128t, 128t+64) = gate rows [64t, 64t+64); rows [128t+64, 128t+128) = up rows
#               [64t, 64t+64)), so that one GEMM tile holds gate and up of the same 64 outputs and
#               the SwiGLU product is taken in the epilogue
#       +5 Wd   [H, I] 16-bit (down_proj)
#
# Names are HF ``MistralModel`` state-dict keys (transformers/models/mistral/modeling_mistral.py).

GATE_UP_BLOCK = 64


def rope_theta_of(hf_config) -> float:
    params = getattr(hf_config, 'rope_parameters', None)
    if params and 'rope_theta' in params:
        return float(params['rope_theta'])
    return float(getattr(hf_config, 'rope_theta', 10000.0))


def mistral_desc(hf_config) -> _native.ModelDesc:
    """Translate a HF ``MistralConfig`` into the C ``B2EModelDesc``."""
    heads = hf_config.num_attention_heads
    head_dim = getattr(hf_config, 'head_dim', None) or hf_config.hidden_size // heads
    if getattr(hf_config, 'hidden_act', 'silu') != 'silu':
        raise NotImplementedError(f'hidden_act={hf_config.hidden_act!r}: only SwiGLU (silu) is built')
    window = getattr(hf_config, 'sliding_window', None)
    return _native.ModelDesc(
        arch=_native.ARCH_MISTRAL,
        num_layers=hf_config.num_hidden_layers,
        hidden=hf_config.hidden_size,
        heads=heads,
        kv_heads=hf_config.num_key_value_heads,
        head_dim=head_dim,
        intermediate=hf_config.intermediate_size,
        vocab=hf_config.vocab_size,
        max_pos=hf_config.max_position_embeddings,
        type_vocab=0,
        eps=float(hf_config.rms_norm_eps),
        rope_theta=rope_theta_of(hf_config),
        sliding_window=int(window) if window else 0,
        reserved=0,
    )


def interleave_gate_up(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """[I,H], [I,H] -> [2I,H] in the block-interleaved row order the SwiGLU epilogue expects."""
    i, h = gate.shape
    if i % GATE_UP_BLOCK:
        raise ValueError(f'intermediate_size {i} must be a multiple of {GATE_UP_BLOCK}')
    g = gate.reshape(i // GATE_UP_BLOCK, 1, GATE_UP_BLOCK, h)
    u = up.reshape(i // GATE_UP_BLOCK, 1, GATE_UP_BLOCK, h)
    return torch.cat([g, u], dim=1).reshape(2 * i, h)


def mistral_weight_list(
    state_dict: Mapping[str, torch.Tensor],
    num_layers: int,
    device: torch.device,
    dtype: torch.dtype = torch.float16,
) -> list[torch.Tensor]:
    """HF MistralModel (or ...ForCausalLM) state dict -> contiguous device tensors in ABI order."""
    sd = {k[6:] if k.startswith('model.') else k: v for k, v in state_dict.items()}

    def f32(key: str) -> torch.Tensor:
        return sd[key].detach().to(device=device, dtype=torch.float32).contiguous()

    def b16(t: torch.Tensor) -> torch.Tensor:
        return to_storage(t, device, dtype)

    out = [f32('embed_tokens.weight'), f32('norm.weight')]
    for layer in range(num_layers):
        p = f'layers.{layer}.'
        qkv = torch.cat([sd[p + f'self_attn.{n}_proj.weight'] for n in ('q', 'k', 'v')])
        out += [
            f32(p + 'input_layernorm.weight'),
            b16(qkv),
            b16(sd[p + 'self_attn.o_proj.weight']),
            f32(p + 'post_attention_layernorm.weight'),
            b16(interleave_gate_up(sd[p + 'mlp.gate_proj.weight'], sd[p + 'mlp.up_proj.weight'])),
            b16(sd[p + 'mlp.down_proj.weight']),
        ]
    return out


def random_mistral_state_dict(hf_config, seed: int = 0, device: torch.device | str = 'cpu',
                              std: float | None = None,
                              dtype: torch.dtype = torch.float32) -> dict[str, torch.Tensor]:
    """Seeded random Mistral weights with HF MistralModel names (no ``model.`` prefix)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    std = getattr(hf_config, 'initializer_range', 0.02) if std is None else std
    h, i = hf_config.hidden_size, hf_config.intermediate_size
    heads, kv = hf_config.num_attention_heads, hf_config.num_key_value_heads
    d = getattr(hf_config, 'head_dim', None) or h // heads

    def normal(*shape: int) -> torch.Tensor:
        return (torch.randn(*shape, generator=gen, device=device, dtype=torch.float32) * std).to(dtype)

    sd: dict[str, torch.Tensor] = {'embed_tokens.weight': normal(hf_config.vocab_size, h)}
    sd['norm.weight'] = (1.0 + normal(h).float()).to(dtype)
    for layer in range(hf_config.num_hidden_layers):
        p = f'layers.{layer}.'
        sd[p + 'self_attn.q_proj.weight'] = normal(heads * d, h)
        sd[p + 'self_attn.k_proj.weight'] = normal(kv * d, h)
        sd[p + 'self_attn.v_proj.weight'] = normal(kv * d, h)
        sd[p + 'self_attn.o_proj.weight'] = normal(h, heads * d)
        sd[p + 'mlp.gate_proj.weight'] = normal(i, h)
        sd[p + 'mlp.up_proj.weight'] = normal(i, h)
        sd[p + 'mlp.down_proj.weight'] = normal(h, i)
        sd[p + 'input_layernorm.weight'] = (1.0 + normal(h).float()).to(dtype)
        sd[p + 'post_attention_layernorm.weight'] = (1.0 + normal(h).float()).to(dtype)
    return sd


# ------------------------------------------------------------------------------ ModernBERT
# ModernBERT (``B2E_ARCH_MODERNBERT``), 5 + 8*L device tensors (HF ``ModernBertModel`` names, ``model.`` prefix
# stripped; transformers/models/modernbert/modeling_modernbert.py):
#
#     0 embeddings.tok_embeddings [V,H] f32      1 embeddings.norm.weight   2 embeddings.norm.bias
#     3 final_norm.weight                         4 final_norm.bias           (f32 [H]; absent biases -> zeros)
#     per layer l, base = 5 + 8*l (pre-LayerNorm blocks; layer 0 has no attn_norm: slots hold ones / zeros):
#       +0 attn_norm.weight   +1 attn_norm.bias
#       +2 attn.Wqkv [3H,H] 16-bit (rows: q | k | v, heads of 64)      +3 attn.Wo [H,H] 16-bit
#       +4 mlp_norm.weight    +5 mlp_norm.bias
#       +6 mlp.Wi [2I,H] 16-bit: the ``input`` half (rows [0,I), the one that goes through GELU) and the ``gate``
#          half (rows [I,2I)) interleaved in blocks of 64 rows, so that the GeGLU product is taken in the GEMM
#          epilogue (``interleave_gate_up(input, gate)``)
#       +7 mlp.Wo [H,I] 16-bit


def modernbert_layer_pattern(hf_config) -> int:
    """``global_every``: layer l is a full-attention layer iff l % global_every == 0 (the published checkpoints:
    3).  Other ``layer_types`` layouts are not built."""
    types = list(hf_config.layer_types)
    for every in range(1, len(types) + 1):
        if all((t == 'full_attention') == (i % every == 0) for i, t in enumerate(types)):
            return every
    raise NotImplementedError(f'layer_types {types} is not "full attention every n-th layer"')


def modernbert_padded_intermediate(intermediate_size: int) -> int:
    """The gated GEMM epilogue pairs 128 input with 128 gate columns: intermediate_size is zero-padded to the
    next multiple of 128 (ModernBERT-large: 2624 -> 2688; gelu(0) * 0 = 0 feeds zero columns of mlp.Wo)."""
    return (intermediate_size + 127) // 128 * 128


def modernbert_desc(hf_config) -> _native.ModelDesc:
    """Translate a HF ``ModernBertConfig`` into the C ``B2EModelDesc``; reject what is not built."""
    if getattr(hf_config, 'hidden_activation', 'gelu') != 'gelu':
        raise NotImplementedError(f'hidden_activation={hf_config.hidden_activation!r}: only erf-GELU is built')
    if getattr(hf_config, 'attention_bias', False) or getattr(hf_config, 'mlp_bias', False):
        raise NotImplementedError('ModernBERT checkpoints with Linear biases are not built')
    heads = hf_config.num_attention_heads
    params = hf_config.rope_parameters
    for kind in ('full_attention', 'sliding_attention'):
        if params[kind].get('rope_type', 'default') != 'default':
            raise NotImplementedError(f'rope_type {params[kind]["rope_type"]!r} is not built')
    return _native.ModelDesc(
        arch=_native.ARCH_MODERNBERT,
        num_layers=hf_config.num_hidden_layers,
        hidden=hf_config.hidden_size,
        heads=heads,
        kv_heads=heads,
        head_dim=hf_config.hidden_size // heads,
        intermediate=modernbert_padded_intermediate(hf_config.intermediate_size),
        vocab=hf_config.vocab_size,
        max_pos=hf_config.max_position_embeddings,
        type_vocab=0,
        eps=float(hf_config.norm_eps),
        rope_theta=float(params['full_attention']['rope_theta']),
        sliding_window=int(hf_config.sliding_window),     # = local_attention // 2: |i - j| <= sliding_window
        reserved=0,
        rope_theta_local=float(params['sliding_attention']['rope_theta']),
        global_every=modernbert_layer_pattern(hf_config),
    )


def modernbert_weight_list(
    state_dict: Mapping[str, torch.Tensor],
    num_layers: int,
    device: torch.device,
    dtype: torch.dtype = torch.bfloat16,
) -> list[torch.Tensor]:
    """HF ModernBertModel (or ...ForMaskedLM) state dict -> contiguous device tensors in ABI order."""
    sd = {k[6:] if k.startswith('model.') else k: v for k, v in state_dict.items()}
    hidden = sd['embeddings.norm.weight'].shape[0]

    def f32(key: str, default: float | None = None) -> torch.Tensor:
        if key not in sd:
            if default is None:
                raise KeyError(key)
            return torch.full((hidden,), default, dtype=torch.float32, device=device)
        return sd[key].detach().to(device=device, dtype=torch.float32).contiguous()

    def b16(t: torch.Tensor) -> torch.Tensor:
        return to_storage(t, device, dtype)

    out = [
        f32('embeddings.tok_embeddings.weight'),
        f32('embeddings.norm.weight'), f32('embeddings.norm.bias', 0.0),
        f32('final_norm.weight'), f32('final_norm.bias', 0.0),
    ]
    for layer in range(num_layers):
        p = f'layers.{layer}.'
        wi = sd[p + 'mlp.Wi.weight'].detach().to(device=device, dtype=torch.float32)
        wo_mlp = sd[p + 'mlp.Wo.weight'].detach().to(device=device, dtype=torch.float32)
        inter = wi.shape[0] // 2
        pad = modernbert_padded_intermediate(inter) - inter
        w_in, w_gate = wi[:inter], wi[inter:]
        if pad:
            zeros = torch.zeros((pad, wi.shape[1]), dtype=wi.dtype, device=wi.device)
            w_in, w_gate = torch.cat([w_in, zeros]), torch.cat([w_gate, zeros])
            wo_mlp = torch.cat([wo_mlp, torch.zeros((wo_mlp.shape[0], pad), dtype=wo_mlp.dtype, device=device)], dim=1)
        out += [
            f32(p + 'attn_norm.weight', 1.0), f32(p + 'attn_norm.bias', 0.0),   # layer 0: Identity (unused)
            b16(sd[p + 'attn.Wqkv.weight']),
            b16(sd[p + 'attn.Wo.weight']),
            f32(p + 'mlp_norm.weight'), f32(p + 'mlp_norm.bias', 0.0),
            b16(interleave_gate_up(w_in, w_gate)),
            b16(wo_mlp),
        ]
    return out


def random_modernbert_state_dict(hf_config, seed: int = 0, device: torch.device | str = 'cpu',
                                 std: float | None = None) -> dict[str, torch.Tensor]:
    """Seeded random ModernBERT weights with HF ModernBertModel names (no biases, as the published models)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    std = getattr(hf_config, 'initializer_range', 0.02) if std is None else std
    h, i = hf_config.hidden_size, hf_config.intermediate_size

    def normal(*shape: int) -> torch.Tensor:
        return torch.randn(*shape, generator=gen, device=device, dtype=torch.float32) * std

    sd: dict[str, torch.Tensor] = {'embeddings.tok_embeddings.weight': normal(hf_config.vocab_size, h)}
    sd['embeddings.norm.weight'] = 1.0 + normal(h)
    sd['final_norm.weight'] = 1.0 + normal(h)
    for layer in range(hf_config.num_hidden_layers):
        p = f'layers.{layer}.'
        if layer > 0:
            sd[p + 'attn_norm.weight'] = 1.0 + normal(h)
        sd[p + 'attn.Wqkv.weight'] = normal(3 * h, h)
        sd[p + 'attn.Wo.weight'] = normal(h, h)
        sd[p + 'mlp_norm.weight'] = 1.0 + normal(h)
        sd[p + 'mlp.Wi.weight'] = normal(2 * i, h)
        sd[p + 'mlp.Wo.weight'] = normal(h, i)
    return sd
