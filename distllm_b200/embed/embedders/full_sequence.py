"""Full-sequence embedder: the batch loop of the hot path.

Replaces distllm/embed/embedders/full_sequence.py:20-80.  When the encoder exposes the fused native
entry point (``encode_pooled``) and the pooler names a native epilogue (``native_pool_kind``) each
batch is ONE call into libb2e: forward pass, pooling and optional L2 normalisation, written
straight into a device-resident ``[N,H]`` matrix.  The reference's per-batch ``.cpu()`` sync
(full_sequence.py:75) becomes a single device->host copy after the loop.  Batches are consumed in
dataloader order with the dataloader's batch size, so the batch-dependent mean-pool quirk sees the
same batch composition as the reference.
"""

from __future__ import annotations

from typing import Literal

import numpy as np
import torch
from pydantic import Field
from torch.utils.data import DataLoader
from tqdm import tqdm

from distllm_b200 import _native
from distllm_b200.embed.embedders.base import EmbedderResult
from distllm_b200.embed.encoders.base import Encoder
from distllm_b200.embed.poolers.base import Pooler
from distllm_b200.utils import BaseConfig


def prefetched(dataloader, depth: int = 2, device: torch.device | None = None):
    """Iterate ``dataloader`` one step ahead on a background thread.

    With ``num_workers == 0`` the tokenizer runs inside ``next(iterator)``; the Rust backend releases
    the GIL, so producing batch i+1 here overlaps the main thread's copy and launches of batch i (the
    GPU work itself is asynchronous either way).  DataLoaders with worker processes already prefetch and
    are passed through untouched.  Order and exceptions are preserved.  ``device``: the CUDA device the
    consumer works on; the producer thread makes it current before it pins memory (DataLoader's own pin
    thread does the same), else pinning would initialise a context on device 0.
    """
    if getattr(dataloader, 'num_workers', 1) != 0 or depth <= 0:
        yield from dataloader
        return
    import queue
    import threading

    done = object()
    box: queue.Queue = queue.Queue(maxsize=depth)
    stop = threading.Event()

    def put(item) -> bool:
        """Stop-aware put: gives up (False) once the consumer has gone away, so the thread never blocks
        forever on a full queue holding the DataLoader iterator and its pinned batches."""
        while not stop.is_set():
            try:
                box.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def produce() -> None:
        try:
            if device is not None and device.type == 'cuda' and torch.cuda.is_available():
                torch.cuda.set_device(device)
            for item in dataloader:
                if not put(item):
                    return
            put(done)
        except BaseException as exc:  # noqa: BLE001  handed to the consumer
            put(exc)

    worker = threading.Thread(target=produce, name='b2e-host-feed', daemon=True)
    worker.start()
    try:
        while True:
            item = box.get()
            if item is done:
                return
            if isinstance(item, BaseException):
                raise item
            yield item
    finally:
        stop.set()
        # drain so that a producer blocked in put() sees the flag promptly, then wait for it to leave
        try:
            while True:
                box.get_nowait()
        except queue.Empty:
            pass
        worker.join(timeout=5.0)


def _fused_kind(encoder: Encoder, pooler: Pooler) -> int | None:
    if not hasattr(encoder, 'encode_pooled'):
        return None
    return getattr(pooler, 'native_pool_kind', None)


@torch.no_grad()
def compute_embeddings_device(
    dataloader: DataLoader,
    encoder: Encoder,
    pooler: Pooler,
    normalize: bool = False,
    progress: bool = True,
) -> torch.Tensor:
    """Pooled embeddings for every row of ``dataloader.dataset`` as a device fp32 ``[N,H]`` tensor."""
    num_embeddings = len(dataloader.dataset)
    out = torch.empty((num_embeddings, encoder.embedding_size), dtype=torch.float32,
                      device=encoder.device)
    kind = _fused_kind(encoder, pooler)
    idx = 0
    for batch in tqdm(prefetched(dataloader, device=encoder.device), total=len(dataloader),
                      disable=not progress):
        inputs = batch.to(encoder.device, non_blocking=True)
        batch_size = inputs['attention_mask'].shape[0]
        if kind is not None:
            encoder.encode_pooled(inputs, kind, normalize, out=out[idx : idx + batch_size])
        else:
            hidden = encoder.encode(inputs)
            pooled = pooler.pool(hidden, inputs['attention_mask']).to(torch.float32)
            if normalize:
                pooled = _native.l2_normalize_(pooled.contiguous())
            out[idx : idx + batch_size] = pooled
        idx += batch_size
    return out


def compute_embeddings_pair(
    dataloader: DataLoader,
    encoder: Encoder,
    pooler: Pooler,
    normalize: bool = False,
) -> tuple[torch.Tensor, np.ndarray]:
    """(device fp32 ``[N,H]``, host ``[N,H]`` array in ``encoder.dtype``): one pass, one D2H copy."""
    device_out = compute_embeddings_device(dataloader, encoder, pooler, normalize)
    return device_out, device_out.to(encoder.dtype).cpu().numpy()


def compute_embeddings(
    dataloader: DataLoader,
    encoder: Encoder,
    pooler: Pooler,
    normalize: bool = False,
) -> np.ndarray:
    """Host ``[N,H]`` array in ``encoder.dtype`` (the reference's return contract, :80)."""
    return compute_embeddings_pair(dataloader, encoder, pooler, normalize)[1]


class FullSequenceEmbedderConfig(BaseConfig):
    """Configuration for the full sequence embedder."""

    name: Literal['full_sequence'] = 'full_sequence'  # type: ignore[assignment]
    normalize_embeddings: bool = Field(
        False,
        description='Whether to return normalized the embeddings.',
    )


class FullSequenceEmbedder:
    """One pooled embedding per input sequence."""

    def __init__(self, config: FullSequenceEmbedderConfig) -> None:
        self.config = config

    def embed(self, dataloader: DataLoader, encoder: Encoder, pooler: Pooler) -> EmbedderResult:
        device_rows, embeddings = compute_embeddings_pair(
            dataloader=dataloader,
            encoder=encoder,
            pooler=pooler,
            normalize=self.config.normalize_embeddings,
        )
        return EmbedderResult(
            embeddings=embeddings,
            text=dataloader.dataset.data,
            metadata=dataloader.dataset.metadata,
            device_embeddings=device_rows,
        )
