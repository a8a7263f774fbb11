"""``auto`` encoder: HuggingFace checkpoint in, native sm_100a forward pass out.

Drop-in for distllm/embed/encoders/auto.py:15-138 -- same config fields and defaults, same
properties, ``encode`` returns the last hidden state ``[B,S,H]``.  transformers is used only to read
the checkpoint and to build the tokenizer; the forward pass is libb2e (fp16 tensor-core GEMMs with
fp32 accumulation, fp32 LayerNorm/softmax statistics).  There is no eager/CPU fallback: an
architecture that is not built raises.
"""

from __future__ import annotations

from typing import Literal
from typing import Optional

import torch
from transformers import BatchEncoding
from transformers import PreTrainedTokenizer

from distllm_b200.embed.encoders.native import NativeBertEncoder
from distllm_b200.embed.encoders.native import NativeMistralEncoder
from distllm_b200.embed.encoders.native import NativeModernBertEncoder
from distllm_b200.utils import BaseConfig

# HF model_type -> native forward pass (BERT: post-LN encoder; Mistral: pre-RMSNorm decoder blocks
# with rotary, grouped-query causal attention and SwiGLU, used as an encoder by the embedding models;
# ModernBERT: pre-LN encoder with rotary, alternating full / sliding-window attention and GeGLU --
# examples/embed/workstation/modernbert_semchunk.yaml:16-17)
_NATIVE_BY_MODEL_TYPE = {'bert': NativeBertEncoder, 'mistral': NativeMistralEncoder,
                         'modernbert': NativeModernBertEncoder}
_SUPPORTED_MODEL_TYPES = tuple(_NATIVE_BY_MODEL_TYPE)


class AutoEncoderConfig(BaseConfig):
    """Config for the AutoModel-compatible encoder (fields as in the reference, auto.py:15-31)."""

    name: Literal['auto'] = 'auto'  # type: ignore[assignment]
    # The model id
    pretrained_model_name_or_path: str
    # Optional tokenizer
    tokenizer_name: Optional[str] = None  # noqa: UP007
    # Report/return half precision (fp16) embeddings
    half_precision: bool = False
    # Kept for compatibility: inference is always in eval mode here
    eval_mode: bool = True
    # Kept for compatibility: there is no tracing compiler in this path
    compile_model: bool = False
    # NF4 (bitsandbytes) weight quantisation, emulated at load time: the GEMMs run on dequant(quant(W))
    quantization: bool = True


class AutoEncoder:
    """Encoder for HF checkpoints of the BERT, ModernBERT and Mistral families on the native kernels."""

    def __init__(self, config: AutoEncoderConfig):
        from transformers import AutoConfig
        from transformers import AutoModel
        from transformers import AutoTokenizer

        hf_config = AutoConfig.from_pretrained(config.pretrained_model_name_or_path)
        if hf_config.model_type not in _SUPPORTED_MODEL_TYPES:
            raise NotImplementedError(
                f'model_type={hf_config.model_type!r} has no native sm_100a forward pass yet '
                f'(built: {_SUPPORTED_MODEL_TYPES}); there is no eager fallback.',
            )
        _NATIVE_BY_MODEL_TYPE[hf_config.model_type].validate(hf_config)   # before any weight is loaded
        model = AutoModel.from_pretrained(config.pretrained_model_name_or_path)
        tokenizer = AutoTokenizer.from_pretrained(
            config.tokenizer_name or config.pretrained_model_name_or_path,
        )
        # proper truncation, as auto.py:74
        tokenizer.model_max_length = hf_config.max_position_embeddings

        self.config = config
        state_dict = model.state_dict()
        if config.quantization:
            # the reference's default (auto.py:44-56): every nn.Linear weight goes through 4-bit NormalFloat with
            # double quantisation and is dequantised in front of each matmul -- what the GEMMs see is
            # dequant(quant(W)).  That tensor is computed once here (embed/encoders/nf4.py restates bitsandbytes'
            # published algorithm; the 4-bit STORAGE is not reproduced, the arithmetic is).
            from distllm_b200.embed.encoders.nf4 import quantize_state_dict_nf4

            state_dict = quantize_state_dict_nf4(state_dict, device='cuda' if torch.cuda.is_available() else None)
        self._native = _NATIVE_BY_MODEL_TYPE[hf_config.model_type](hf_config, state_dict)
        del model, state_dict
        self._tokenizer = tokenizer
        self._dtype = torch.float16 if config.half_precision else torch.float32

    @classmethod
    def from_native(cls, native: NativeBertEncoder, tokenizer: PreTrainedTokenizer | None = None,
                    half_precision: bool = False) -> 'AutoEncoder':
        """Wrap an already-built native encoder (synthetic weights, tests, benchmarks)."""
        self = cls.__new__(cls)
        self.config = None
        self._native = native
        self._tokenizer = tokenizer
        self._dtype = torch.float16 if half_precision else torch.float32
        return self

    @property
    def dtype(self) -> torch.dtype:
        return self._dtype

    @property
    def device(self) -> torch.device:
        return self._native.device

    @property
    def embedding_size(self) -> int:
        return self._native.hidden_size

    @property
    def tokenizer(self) -> PreTrainedTokenizer:
        return self._tokenizer

    @property
    def native(self) -> NativeBertEncoder:
        return self._native

    def encode(self, batch_encoding: BatchEncoding) -> torch.Tensor:
        """Last hidden state ``[B,S,H]`` in ``self.dtype`` (auto.py:119-138)."""
        hidden = self._native.encode(
            batch_encoding['input_ids'],
            batch_encoding['attention_mask'],
            batch_encoding.get('token_type_ids'),
            out_dtype=torch.float32,
        )
        return hidden if self._dtype == torch.float32 else hidden.to(self._dtype)

    def encode_pooled(self, batch_encoding: BatchEncoding, pool_kind: int, normalize: bool,
                      out: torch.Tensor | None = None) -> torch.Tensor:
        """Fused encode + pool (+ normalise) -> fp32 ``[B,H]`` (used by the native embedders)."""
        return self._native.encode_pooled(
            batch_encoding['input_ids'],
            batch_encoding['attention_mask'],
            batch_encoding.get('token_type_ids'),
            pool_kind,
            normalize,
            out=out,
        )
