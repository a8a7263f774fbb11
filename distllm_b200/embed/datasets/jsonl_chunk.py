"""jsonl documents -> sentence buffers for semantic chunking.

Follows distllm/embed/datasets/jsonl_chunk.py:24-179: split each document into sentences keeping
the whitespace that follows a sentence attached to it, build one buffer per sentence from the
sentences within ``buffer_size`` of it, carry a copy of the document metadata plus the sentence on
every row, and drop buffers of ``min_buffer_length`` characters or fewer.
"""

from __future__ import annotations

import re
import warnings
from pathlib import Path
from typing import Any
from typing import Callable
from typing import Literal

from pydantic import Field
from torch.utils.data import DataLoader

from distllm_b200.embed.datasets.jsonl import read_jsonl
from distllm_b200.embed.datasets.utils import InMemoryDataset
from distllm_b200.embed.datasets.utils import SentenceTokenCache
from distllm_b200.embed.datasets.utils import LoaderConfig
from distllm_b200.embed.datasets.utils import make_dataloader
from distllm_b200.embed.encoders.base import Encoder
from distllm_b200.utils import BaseConfig

# sentence end = . ! or ? (plus closing quotes/brackets), whitespace, then an upper-case/digit start
_BOUNDARY = re.compile(r'[.!?]["\')\]]*\s+(?=[A-Z0-9"\'(\[])')


def _regex_spans(text: str) -> list[tuple[int, int]]:
    """Dependency-free stand-in for Punkt ``span_tokenize`` (used only when nltk is absent)."""
    spans = []
    start = 0
    for m in _BOUNDARY.finditer(text):
        end = m.end()
        stop = m.start() + len(m.group().rstrip())
        spans.append((start, stop))
        start = end
    if start < len(text):
        spans.append((start, len(text.rstrip()) if text.rstrip() else len(text)))
    return [(s, e) for s, e in spans if e > s]


def _span_tokenizer(allow_regex_fallback: bool = True) -> Callable[[str], list[tuple[int, int]]]:
    """Punkt ``span_tokenize`` (what the reference uses, jsonl_chunk.py:26-28).  Without nltk the regex
    stand-in is used ONLY when allowed, and never silently: its sentence boundaries differ from Punkt's
    on abbreviations and the like, so buffers, chunks and embeddings then differ from the reference."""
    try:
        import nltk

        punkt = nltk.tokenize.PunktSentenceTokenizer()
        return lambda text: list(punkt.span_tokenize(text))
    except ImportError:
        if not allow_regex_fallback:
            raise ImportError(
                'nltk is required for the reference sentence splitter (PunktSentenceTokenizer); '
                'install it or set sentence_splitter="regex" in the jsonl_chunk dataset config') from None
        import warnings

        warnings.warn(
            'nltk is not installed: jsonl_chunk splits sentences with a regex stand-in whose boundaries '
            'differ from the reference (Punkt) on abbreviations; chunk texts may differ from distllm. '
            'Set sentence_splitter="regex" to silence this, or install nltk.',
            RuntimeWarning, stacklevel=3)
        return _regex_spans


def split_by_sentence_tokenizer(splitter: str = 'auto') -> Callable[[str], list[str]]:
    """Sentence splitter whose pieces concatenate back to the original text (minus leading junk):
    piece ``i`` runs from the start of sentence ``i`` to the start of sentence ``i+1``.

    ``splitter``: 'punkt' (nltk required, the reference), 'regex' (explicit opt-in to the stand-in),
    'auto' (punkt when nltk is importable, else the stand-in with a RuntimeWarning)."""
    if splitter == 'regex':
        spans_of = _regex_spans
    elif splitter == 'punkt':
        spans_of = _span_tokenizer(allow_regex_fallback=False)
    elif splitter == 'auto':
        spans_of = _span_tokenizer()
    else:
        raise ValueError(f"sentence_splitter must be 'auto', 'punkt' or 'regex', got {splitter!r}")

    def split(text: str) -> list[str]:
        starts = [s for s, _ in spans_of(text)]
        stops = starts[1:] + [len(text)]
        return [text[a:b] for a, b in zip(starts, stops)]

    return split


def sentences_to_buffers(split: list[str], buffer_size: int) -> list[str]:
    """Buffer ``i`` = sentences ``[i - buffer_size, i + buffer_size]`` joined (clipped at the ends)."""
    n = len(split)
    return [''.join(split[max(0, i - buffer_size) : min(n, i + buffer_size + 1)]) for i in range(n)]


class JsonlChunkDatasetConfig(LoaderConfig):
    name: Literal['jsonl_chunk'] = 'jsonl_chunk'  # type: ignore[assignment]
    text_field: str = 'text'   # which key of each json row holds the text
    # tokenise every sentence once and assemble buffer / chunk ids from the cached pieces (BERT-style tokenizers
    # only, identical ids; False = tokenise every buffer and chunk text like the reference)
    sentence_token_cache: bool = True
    min_buffer_length: int = Field(
        default=750,
        description='Buffers with this many characters or fewer are filtered out '
        '(removes citations and other fragments).',
    )
    buffer_size: int = Field(
        default=1,
        description='Number of neighbouring sentences on each side grouped with a sentence '
        'when evaluating semantic similarity.',
    )
    # not in the reference (it hard-requires nltk): which sentence splitter to use, see
    # split_by_sentence_tokenizer
    sentence_splitter: Literal['auto', 'punkt', 'regex'] = 'auto'


class JsonlChunkDataset:
    def __init__(self, config: JsonlChunkDatasetConfig):
        self.config = config
        self.splitter = split_by_sentence_tokenizer(config.sentence_splitter)

    def get_dataloader(self, data_file: Path, encoder: Encoder) -> DataLoader:
        rows: list[dict[str, Any]] = read_jsonl(data_file)
        texts = [row.pop(self.config.text_field) for row in rows]
        # whatever is left of each row is that document's metadata
        if not rows or any(not row for row in rows):
            raise ValueError('Metadata is empty. Please check the jsonl file.')

        buffers: list[str] = []
        metadatas: list[dict[str, Any]] = []
        all_sentences: list[str] = []          # every sentence of the file, in order
        parts: list[tuple[int, ...]] = []      # buffer -> the sentences it is joined from
        n_buf = self.config.buffer_size
        for doc_meta, text in zip(rows, texts):
            sentences = self.splitter(text)
            base, n = len(all_sentences), len(sentences)
            all_sentences.extend(sentences)
            buffers.extend(sentences_to_buffers(sentences, n_buf))
            parts.extend(tuple(range(base + max(0, i - n_buf), base + min(n, i + n_buf + 1))) for i in range(n))
            metadatas.extend({**doc_meta, 'sentence': sentence} for sentence in sentences)

        keep = [i for i, buf in enumerate(buffers) if len(buf) > self.config.min_buffer_length]
        # every sentence is tokenised once instead of once per buffer that contains it (2 * buffer_size + 1 times)
        cache = SentenceTokenCache.build(encoder.tokenizer, all_sentences) if self.config.sentence_token_cache else None
        dataset = InMemoryDataset([buffers[i] for i in keep], [metadatas[i] for i in keep],
                                  parts=[parts[i] for i in keep], token_cache=cache, sentence_index=keep)
        if cache is not None and len(dataset):
            sample = sorted({0, len(dataset) // 3, (2 * len(dataset)) // 3, len(dataset) - 1})
            if not cache.agrees_with(encoder.tokenizer, [dataset[i] for i in sample]):
                warnings.warn('sentence token cache disagrees with the tokenizer on this file: tokenising every '
                              'buffer directly', RuntimeWarning, stacklevel=2)
                dataset.parts = dataset.token_cache = None
        return make_dataloader(self.config, dataset, encoder.tokenizer)
