"""Host-side mirror of the reference's plugin surface: names, defaults, errors, file formats."""

from __future__ import annotations

import json

import numpy as np
import pytest

from distllm_b200.embed import get_dataset
from distllm_b200.embed import get_embedder
from distllm_b200.embed import get_pooler
from distllm_b200.embed import get_writer
from distllm_b200.embed.datasets.fasta import read_fasta
from distllm_b200.embed.datasets.fasta import write_fasta
from distllm_b200.embed.datasets.fasta import Sequence
from distllm_b200.embed.datasets.jsonl_chunk import sentences_to_buffers
from distllm_b200.embed.datasets.jsonl_chunk import split_by_sentence_tokenizer
from distllm_b200.embed.embedders.base import EmbedderResult
from distllm_b200.embed.embedders.semantic_chunk import build_chunks
from distllm_b200.embed.embedders.semantic_chunk import document_ranges
from distllm_b200.registry import RegistrySingleton
from distllm_b200.registry import registry
from distllm_b200.sharding import shard_range
from distllm_b200.timer import TimeLogger
from distllm_b200.timer import Timer
from distllm_b200.utils import BaseConfig
from distllm_b200.utils import batch_data


def test_strategy_names_and_defaults():
    import distllm_b200.embed.datasets as ds
    import distllm_b200.embed.embedders as em
    import distllm_b200.embed.encoders as en
    import distllm_b200.embed.poolers as po
    import distllm_b200.embed.writers as wr

    assert set(po.STRATEGIES) == {'mean', 'last_token'}
    assert set(em.STRATEGIES) == {'full_sequence', 'semantic_chunk'}
    assert set(wr.STRATEGIES) == {'huggingface', 'numpy'}
    assert {'jsonl', 'jsonl_chunk', 'fasta', 'sequence_per_line'} <= set(ds.STRATEGIES)
    assert 'auto' in en.STRATEGIES

    sc = get_embedder({'name': 'semantic_chunk'}).config
    assert (sc.breakpoint_percentile_threshold, sc.chunk_batch_size, sc.min_chunk_length,
            sc.normalize_embeddings) == (90, 8, 750, False)
    jc = get_dataset({'name': 'jsonl_chunk'}).config
    assert (jc.text_field, jc.num_data_workers, jc.batch_size, jc.pin_memory, jc.min_buffer_length,
            jc.buffer_size) == ('text', 4, 8, True, 750, 1)
    auto_cfg = en.AutoEncoderConfig(pretrained_model_name_or_path='x')
    assert (auto_cfg.half_precision, auto_cfg.eval_mode, auto_cfg.compile_model,
            auto_cfg.quantization, auto_cfg.tokenizer_name) == (False, True, False, True, None)
    assert get_embedder({'name': 'full_sequence'}).config.normalize_embeddings is False
    assert get_writer({'name': 'huggingface'}).config.num_proc is None


@pytest.mark.parametrize('factory', [get_pooler, get_embedder, get_writer, get_dataset])
def test_unknown_name_raises_value_error(factory):
    with pytest.raises(ValueError, match='Unknown .* name: nope'):
        factory({'name': 'nope'})


def test_unknown_encoder_raises_value_error():
    from distllm_b200.embed import get_encoder

    with pytest.raises(ValueError, match='Unknown encoder name'):
        get_encoder({'name': 'nope'})


def test_config_yaml_json_roundtrip(tmp_path):
    from distllm_b200.distributed_embedding import Config

    cfg = Config(
        input_dir=tmp_path, output_dir=tmp_path / 'out', glob_patterns=['*.jsonl'],
        dataset_config={'name': 'jsonl_chunk', 'batch_size': 512, 'buffer_size': 4},
        encoder_config={'name': 'auto', 'pretrained_model_name_or_path': 'pritamdeka/S-PubMedBert-MS-MARCO',
                        'quantization': False},
        pooler_config={'name': 'mean'},
        embedder_config={'name': 'semantic_chunk', 'chunk_batch_size': 512},
        writer_config={'name': 'numpy'},
        compute_config={'name': 'workstation', 'available_accelerators': 8},
    )
    cfg.write_yaml(tmp_path / 'c.yaml')
    back = Config.from_yaml(tmp_path / 'c.yaml')
    assert back == cfg
    assert type(back.dataset_config).__name__ == 'JsonlChunkDatasetConfig'
    assert type(back.embedder_config).__name__ == 'SemanticChunkEmbedderConfig'
    cfg.write_json(tmp_path / 'c.json')
    assert json.loads((tmp_path / 'c.json').read_text())['pooler_config']['name'] == 'mean'

    class Tiny(BaseConfig):
        x: int = 3

    Tiny(x=5).write_json(tmp_path / 't.json')
    assert Tiny.from_json(tmp_path / 't.json').x == 5


def test_batch_data():
    assert batch_data(list(range(7)), 3) == [[0, 1, 2], [3, 4, 5], [6]]
    assert batch_data([], 3) == []
    assert batch_data([1, 2], 5) == [[1, 2]]


def test_timer_lines_parse_back(tmp_path, capsys):
    with Timer('computed-embeddings', 'file.jsonl'):
        pass
    t = Timer('loaded-encoder').start()
    with pytest.raises(RuntimeError):
        _ = t.elapsed_ns
    t.stop()
    out = capsys.readouterr().out
    assert out.count('[timer]') == 2
    log = tmp_path / 'log.txt'
    log.write_text('noise\n' + out)
    stats = TimeLogger().parse_logs(log)
    assert [list(s.tags) for s in stats] == [['computed-embeddings', 'file.jsonl'], ['loaded-encoder']]
    assert float(stats[0].elapsed_s) >= 0.0
    assert float(stats[0].end_unix) >= float(stats[0].start_unix)


def test_registry_warm_start_and_eviction():
    assert RegistrySingleton() is registry
    registry.clear()
    built, closed = [], []

    def factory(**kw):
        built.append(kw)
        return dict(kw)

    registry.register(factory, shutdown_callback=closed.append)
    a = registry.get(factory, name='auto', path='m1')
    assert registry.get(factory, name='auto', path='m1') is a and len(built) == 1
    b = registry.get(factory, name='auto', path='m2')
    assert b is not a and len(built) == 2 and closed == [a]
    with pytest.raises(ValueError, match='not registered'):
        registry.get(lambda: None)
    registry.clear()
    assert closed[-1] is b


def test_sentence_split_is_lossless_and_buffers_window():
    text = 'Alpha beta gamma.  Delta epsilon!\nZeta eta? Theta iota.'
    parts = split_by_sentence_tokenizer()(text)
    assert ''.join(parts) == text
    assert len(parts) == 4 and parts[0] == 'Alpha beta gamma.  '
    bufs = sentences_to_buffers(list('abcde'), 1)
    assert bufs == ['ab', 'abc', 'bcd', 'cde', 'de']
    assert sentences_to_buffers(list('abc'), 4) == ['abc'] * 3
    assert sentences_to_buffers([], 2) == []


def test_jsonl_chunk_rows_and_filter(tmp_path):
    class FakeEncoder:
        tokenizer = staticmethod(lambda batch, **kw: batch)

    sent = 'Sentence number %d has quite a few words in it. '
    doc = ''.join(sent % i for i in range(12))
    f = tmp_path / 'd.jsonl'
    f.write_text('\n'.join(json.dumps({'text': doc, 'path': f'p{i}'}) for i in range(2)))
    ds = get_dataset({'name': 'jsonl_chunk', 'buffer_size': 2, 'min_buffer_length': 150,
                      'num_data_workers': 0, 'pin_memory': False})
    loader = ds.get_dataloader(f, FakeEncoder())
    data, meta = loader.dataset.data, loader.dataset.metadata
    # buffers at the document edges (3 sentences ~ 147 chars) are filtered, the rest kept
    assert len(data) == len(meta) == 2 * 10
    assert all(len(d) > 150 for d in data)
    assert meta[0]['path'] == 'p0' and meta[-1]['path'] == 'p1' and 'sentence' in meta[0]
    assert document_ranges(meta) == [(0, 10), (10, 20)]
    with pytest.raises(ValueError, match='Metadata is empty'):
        f.write_text(json.dumps({'text': doc}))
        ds.get_dataloader(f, FakeEncoder())


def test_fasta_roundtrip(tmp_path):
    f = tmp_path / 'x.fasta'
    write_fasta([Sequence('MKV', 'a b'), Sequence('ACDE', 't2')], f)
    with open(f, 'a') as h:
        h.write('>t3\nAA\nCC\n')
    recs = read_fasta(f)
    assert [(r.tag, r.sequence) for r in recs] == [('a b', 'MKV'), ('t2', 'ACDE'), ('t3', 'AACC')]


def test_numpy_writer_write_and_merge(tmp_path):
    w = get_writer({'name': 'numpy'})
    dirs = []
    for k in range(2):
        d = tmp_path / f'r{k}'
        d.mkdir()
        w.write(d, EmbedderResult(np.full((3, 4), k, np.float32), [f't{k}{i}' for i in range(3)],
                                  [{'path': f'p{k}'}] * 3))
        dirs.append(d)
    out = tmp_path / 'merged'
    out.mkdir()
    w.merge(dirs, out)
    assert np.load(out / 'embeddings.npy').shape == (6, 4)
    assert list(np.load(out / 'text.npy'))[3] == 't10'
    assert np.load(out / 'metadata.npy', allow_pickle=True)[5]['path'] == 'p1'


def test_huggingface_writer_schema(tmp_path):
    from datasets import Dataset

    w = get_writer({'name': 'huggingface'})
    w.write(tmp_path / 'ds', EmbedderResult(np.arange(8, dtype=np.float32).reshape(2, 4), ['a', 'b'],
                                            [{'path': 'p', 'k': 1}, {'path': 'q', 'k': 2}]))
    ds = Dataset.load_from_disk(tmp_path / 'ds')
    assert ds.column_names == ['text', 'embeddings', 'path', 'k']
    assert ds[1]['embeddings'] == [4.0, 5.0, 6.0, 7.0] and ds[1]['path'] == 'q'


def test_build_chunks_matches_reference_groups(semantic_golden):
    """The product's vectorised build_chunks against groups produced by the reference."""
    for k in range(len(semantic_golden['doc_ranges'])):
        for pct in (50, 90, 95):
            want = [tuple(int(v) for v in g) for g in semantic_golden[f'groups/{k}/{pct}']]
            assert build_chunks(semantic_golden[f'dist/{k}'], pct) == want
    assert build_chunks(np.zeros(0), 90) == [(0, 0)]


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 8, 9, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_cli_embed_flags_match_reference():
    import typer

    from distllm_b200.cli import app

    group = typer.main.get_command(app)
    embed = group.commands['embed']
    opts = {o for p in embed.params for o in p.opts}
    want = {'--encoder_name': '-mn', '--pretrained_model_name_or_path': '-m', '--data_path': '-d',
            '--data_extension': '-de', '--output_path': '-o', '--dataset_name': '-dn',
            '--batch_size': '-b', '--chunk_batch_size': '-cb', '--buffer_size': '-bs',
            '--pooler_name': '-pn', '--embedder_name': '-en', '--writer_name': '-wn',
            '--half_precision': '-hp', '--eval_mode': '-em', '--compile_model': '-cm',
            '--quantization': '-q'}
    for long, short in want.items():
        assert long in opts and short in opts, (long, short)
    defaults = {p.name: p.default for p in embed.params}
    assert (defaults['dataset_name'], defaults['batch_size'], defaults['pooler_name'],
            defaults['embedder_name'], defaults['writer_name'], defaults['quantization']) == (
        'jsonl', 1, 'mean', 'full_sequence', 'huggingface', False)
    merge_opts = {o for p in group.commands['merge'].params for o in p.opts}
    assert {'--writer_name', '--num_proc', '--dataset_dir', '--output_dir'} <= merge_opts


# ------------------------------------------------------------------------------ host feed (collator)
def _bert_tokenizer(tmp_path, max_len=32):
    from transformers import BertTokenizerFast

    words = [f'w{i:03d}' for i in range(300)]
    (tmp_path / 'vocab.txt').write_text('\n'.join(['[PAD]', '[UNK]', '[CLS]', '[SEP]', '[MASK]', *words]) + '\n')
    tok = BertTokenizerFast(vocab=str(tmp_path / 'vocab.txt'), do_lower_case=False)
    tok.model_max_length = max_len
    return tok, words


def _llama_like_tokenizer(max_len=24):
    from tokenizers import Tokenizer
    from tokenizers.models import WordLevel
    from tokenizers.pre_tokenizers import Whitespace
    from tokenizers.processors import TemplateProcessing
    from transformers import PreTrainedTokenizerFast

    words = [f'w{i:03d}' for i in range(300)]
    vocab = {t: i for i, t in enumerate(['<pad>', '<s>', '</s>', '<unk>', *words])}
    raw = Tokenizer(WordLevel(vocab, unk_token='<unk>'))
    raw.pre_tokenizer = Whitespace()
    raw.post_processor = TemplateProcessing(single='<s> $A', special_tokens=[('<s>', 1)])
    tok = PreTrainedTokenizerFast(tokenizer_object=raw, pad_token='<pad>', bos_token='<s>', eos_token='</s>',
                                  unk_token='<unk>', model_input_names=['input_ids', 'attention_mask'])
    tok.model_max_length = max_len
    return tok, words


def _assert_same_batch(collator_fast, collator_ref, texts):
    import torch

    a, b = collator_fast(texts), collator_ref(texts)
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].dtype == b[k].dtype == torch.int64
        assert torch.equal(a[k], b[k]), k


def test_fast_collator_equals_the_reference_call_bert(tmp_path):
    """The numpy-padded batch built from the Rust backend's encodings is, tensor for tensor, what
    `tokenizer(batch, padding=True, truncation=True, return_tensors='pt')` returns
    (distllm/embed/datasets/utils.py:43-50): ragged rows, a truncated row, an empty string, unknown
    words, one-row batches."""
    from distllm_b200.embed.datasets.utils import DataCollator

    tok, words = _bert_tokenizer(tmp_path)
    rng = np.random.default_rng(3)
    texts = [' '.join(rng.choice(words, size=n)) for n in (1, 7, 30, 31, 80, 2, 15)] + ['', 'zzz unknown w001']
    fast, ref = DataCollator(tok), DataCollator(tok, fast=False)
    assert fast._fast and not ref._fast
    _assert_same_batch(fast, ref, texts)
    _assert_same_batch(fast, ref, texts[:1])
    _assert_same_batch(fast, ref, [texts[4]])          # only a truncated row
    assert fast(texts)['input_ids'].shape[1] == 32     # truncated to model_max_length
    # the reference call still works after the fast path has reconfigured the backend, and vice versa
    _assert_same_batch(fast, ref, texts[::-1])


@pytest.mark.parametrize('side', ['right', 'left'])
def test_fast_collator_equals_the_reference_call_without_token_types(side):
    from distllm_b200.embed.datasets.utils import DataCollator

    tok, words = _llama_like_tokenizer()
    tok.padding_side = side
    rng = np.random.default_rng(4)
    texts = [' '.join(rng.choice(words, size=n)) for n in (3, 40, 1, 23, 22, 9)]
    fast, ref = DataCollator(tok), DataCollator(tok, fast=False)
    _assert_same_batch(fast, ref, texts)
    assert 'token_type_ids' not in fast(texts)
    mask = fast(texts)['attention_mask']
    assert (mask[:, 0] == 1).all() if side == 'right' else (mask[:, -1] == 1).all()


def test_slow_tokenizers_take_the_reference_call(tmp_path):
    from transformers import EsmTokenizer

    from distllm_b200.embed.datasets.utils import DataCollator

    (tmp_path / 'vocab.txt').write_text('\n'.join(['<cls>', '<pad>', '<eos>', '<unk>', 'L', 'A', 'G', 'V', '<mask>']) + '\n')
    tok = EsmTokenizer(str(tmp_path / 'vocab.txt'))
    coll = DataCollator(tok)
    assert not coll._fast
    out = coll(['LAGV', 'LA'])
    assert out['input_ids'].shape == (2, 6) and out['attention_mask'].sum().item() == 10


# ------------------------------------------------------------------------------ columnar writer
def _reference_rows(result):
    """What distllm/embed/writers/huggingface.py:19-34 + :70 build: one dict per row -> from_list."""
    rows = []
    for idx, (text, emb) in enumerate(zip(result.text, result.embeddings)):
        item = {'text': text, 'embeddings': emb}
        if result.metadata is not None:
            item.update(result.metadata[idx])
        rows.append(item)
    return rows


@pytest.mark.parametrize('dtype', [np.float32, np.float16])
@pytest.mark.parametrize('with_metadata', [True, False])
def test_columnar_huggingface_writer_equals_the_reference_table(tmp_path, dtype, with_metadata):
    """Same features, same column order, same content as the reference's row-wise `Dataset.from_list`,
    including a metadata key that is missing from later rows and one that only later rows carry
    (dropped by from_list, which takes its columns from the first row)."""
    from datasets import Dataset

    rng = np.random.default_rng(1)
    n, h = 37, 24
    emb = rng.standard_normal((n, h)).astype(dtype)
    text = [f'chunk {i}' for i in range(n)]
    meta = None
    if with_metadata:
        meta = [{'path': f'doc{i // 5}', 'idx': i, 'score': i / 7} for i in range(n)]
        del meta[3]['score']          # absent later -> None
        meta[5]['extra'] = 'ignored'  # not in the first row -> no column
    result = EmbedderResult(emb, text, meta)
    w = get_writer({'name': 'huggingface'})
    w.write(tmp_path / 'col', result)
    got = Dataset.load_from_disk(tmp_path / 'col')
    want = Dataset.from_list(_reference_rows(result))
    assert got.features == want.features
    assert got.column_names == want.column_names
    assert got.to_dict() == want.to_dict()
    # and the merged dataset of two such directories reads back in order
    w.write(tmp_path / 'col2', result)
    w.merge([tmp_path / 'col', tmp_path / 'col2'], tmp_path / 'merged')
    merged = Dataset.load_from_disk(tmp_path / 'merged')
    assert len(merged) == 2 * n and merged[n]['text'] == 'chunk 0'


def test_columnar_writer_falls_back_on_untypeable_metadata(tmp_path):
    from datasets import Dataset

    result = EmbedderResult(np.zeros((2, 4), np.float32), ['a', 'b'], [{'k': 1}, {'k': 'one'}])
    w = get_writer({'name': 'huggingface'})
    try:
        want = Dataset.from_list(_reference_rows(result))
    except Exception:  # noqa: BLE001  the reference cannot type this column either
        want = None
    if want is None:
        with pytest.raises(Exception):  # noqa: B017, PT011
            w.write(tmp_path / 'x', result)
    else:
        w.write(tmp_path / 'x', result)
        assert Dataset.load_from_disk(tmp_path / 'x').to_dict() == want.to_dict()


# ------------------------------------------------------------------------------ retrieval (host side)
def test_search_oracle_and_score_filter():
    from datasets.search import BatchedSearchResults

    from distllm_b200.rag.search import ExactIndex
    from distllm_b200.rag.search import filter_search_by_score
    from oracle import search as osearch

    rng = np.random.default_rng(2)
    q = rng.standard_normal((4, 16)).astype(np.float32)
    c = rng.standard_normal((50, 16)).astype(np.float32)
    c[7] = c[3]                                  # a tie: the lower index comes first
    s, i = osearch.topk_inner_product(q, c, 10)
    full = q @ c.T
    for r in range(4):
        assert np.allclose(s[r], np.sort(full[r])[::-1][:10], rtol=1e-6)
        assert set(i[r]) == set(np.argsort(-full[r], kind='stable')[:10])
        pos = {int(v): p for p, v in enumerate(i[r])}
        if 3 in pos and 7 in pos:
            assert pos[3] < pos[7]
    s2, i2 = osearch.topk_inner_product(q, c[:3], 10)
    assert s2.shape == (4, 3) and i2.shape == (4, 3)
    # normalisation: faiss.normalize_L2 semantics, zero rows stay zero, in place
    x = np.array([[3.0, 4.0], [0.0, 0.0]], dtype=np.float32)
    assert ExactIndex.transform(x) is x and np.allclose(x, [[0.6, 0.8], [0.0, 0.0]])
    assert np.allclose(osearch.normalize_l2(np.array([[3.0, 4.0]])), [[0.6, 0.8]])
    res = BatchedSearchResults(total_scores=[[0.9, 0.5, 0.1], [0.2]], total_indices=[[4, 2, 9], [1]])
    assert filter_search_by_score(res, 0.0) is res
    kept = filter_search_by_score(res, 0.5)
    assert kept.total_indices == [[4, 2], []] and kept.total_scores == [[0.9, 0.5], []]


@pytest.mark.skipif(__import__('torch').cuda.is_available(), reason='CPU-only behaviour')
def test_exact_index_has_no_cpu_fallback():
    from distllm_b200 import _native
    from distllm_b200.rag.search import ExactIndex

    with pytest.raises(_native.NativeError, match='no CPU fallback'):
        ExactIndex(np.zeros((4, 128), np.float32))
    with pytest.raises(ValueError, match='embedding matrix or a dataset_dir'):
        ExactIndex()


def test_prefetched_iteration_keeps_order_and_errors():
    import threading

    from distllm_b200.embed.embedders.full_sequence import prefetched

    class Loader:
        num_workers = 0

        def __init__(self, n, fail_at=None):
            self.n, self.fail_at, self.threads = n, fail_at, set()

        def __len__(self):
            return self.n

        def __iter__(self):
            for i in range(self.n):
                self.threads.add(threading.current_thread().name)
                if i == self.fail_at:
                    raise KeyError('boom')
                yield i

    ld = Loader(25)
    assert list(prefetched(ld)) == list(range(25))
    assert ld.threads == {'b2e-host-feed'}                     # produced off the main thread
    with pytest.raises(KeyError, match='boom'):
        list(prefetched(Loader(10, fail_at=4)))
    it = prefetched(Loader(1000))                                # abandoned early: the producer stops
    assert [next(it) for _ in range(3)] == [0, 1, 2]
    it.close()
    # ... and its thread is gone: also when it was blocked on a FULL queue holding the final sentinel
    # (3 items, depth 2, consumer leaves after the first) or an exception
    for loader in (Loader(3), Loader(3, fail_at=2)):
        it = prefetched(loader, depth=2)
        assert next(it) == 0
        it.close()
    assert not [t for t in threading.enumerate() if t.name == 'b2e-host-feed' and t.is_alive()]
    ld4 = Loader(5)
    ld4.num_workers = 4                                          # worker processes: passed through
    assert list(prefetched(ld4)) == list(range(5)) and ld4.threads == {threading.current_thread().name}


def test_retriever_config_roundtrip(tmp_path):
    from distllm_b200.rag import RetrieverConfig

    cfg = RetrieverConfig(
        faiss_config={'dataset_dir': str(tmp_path / 'ds'), 'corpus_dtype': 'bfloat16'},
        encoder_config={'name': 'auto', 'pretrained_model_name_or_path': 'x', 'quantization': False},
        pooler_config={'name': 'mean'},
    )
    assert cfg.batch_size == 4 and cfg.faiss_config.search_algorithm == 'exact'
    cfg.write_yaml(tmp_path / 'r.yaml')
    assert RetrieverConfig.from_yaml(tmp_path / 'r.yaml') == cfg
    with pytest.raises(Exception):  # noqa: B017, PT011  only the exact search exists
        RetrieverConfig(faiss_config={'search_algorithm': 'hnsw'}, encoder_config={}, pooler_config={})


# ------------------------------------------------------------------------------ kernel host-math mirrors
def _at4_ranges(t: int, nq: int, window: int, kv_chunks: int):
    """Python mirror of at4_decode (distllm_b200/csrc/attention4.cuh): chunk range a query tile visits."""
    if t >= nq:
        return 0, 0
    hi = min(kv_chunks, 2 * t + 2)
    lo = max(0, 128 * t - window + 1) // 64 if window > 0 else 0
    if lo >= hi:
        lo = hi - 1
    return lo, hi


def _at4_edge(t: int, j: int, window: int) -> bool:
    key0 = 64 * j
    return key0 + 63 > 128 * t or (window > 0 and 128 * t + 127 - key0 >= window)


@pytest.mark.parametrize('window', [0, 1, 64, 65, 80, 128, 300, 4096])
def test_causal_window_chunk_ranges_cover_every_visible_key(window):
    """The chunk range a query tile walks contains every chunk with a visible key of any of its rows, and
    a chunk classified as interior (unmasked fast path) is fully visible to all 128 rows of the tile."""
    for s in (1, 63, 64, 65, 128, 129, 400, 1100):
        for valid in {1, s // 2 + 1, s}:                      # right-padded length
            nq = (s + 127) // 128
            kv_chunks = (valid + 63) // 64
            for t in range(nq):
                lo, hi = _at4_ranges(t, nq, window, kv_chunks)
                assert 0 <= lo < hi <= kv_chunks
                rows = range(128 * t, min(128 * (t + 1), s))
                for i in rows:
                    first = max(0, i - window + 1) if window else 0
                    last = min(i, valid - 1)
                    if first <= last:                          # the row sees attended keys
                        assert lo <= first // 64 and last // 64 < hi, (s, valid, t, i)
                for j in range(lo, hi):
                    if not _at4_edge(t, j, window):
                        for i in (128 * t, 128 * t + 127):
                            assert 64 * j + 63 <= i and (not window or i - 64 * j < window)


def test_huggingface_dataset_strategy(tmp_path):
    """`huggingface` dataset strategy (ref embed/datasets/huggingface.py:18-83): text column + optional
    metadata columns of a save_to_disk directory."""
    import datasets

    class FakeEncoder:
        tokenizer = staticmethod(lambda batch, **kw: batch)

    rows = {'text': [f'row {i}' for i in range(5)], 'path': [f'p{i}' for i in range(5)], 'n': list(range(5))}
    datasets.Dataset.from_dict(rows).save_to_disk(str(tmp_path / 'ds'))
    cfg = get_dataset({'name': 'huggingface'}).config
    assert (cfg.text_field, cfg.metadata_fields, cfg.num_data_workers, cfg.batch_size, cfg.pin_memory) == (
        'text', [], 4, 8, True)
    plain = get_dataset({'name': 'huggingface', 'num_data_workers': 0, 'pin_memory': False})
    loader = plain.get_dataloader(tmp_path / 'ds', FakeEncoder())
    assert loader.dataset.data == rows['text'] and loader.dataset.metadata is None
    with_meta = get_dataset({'name': 'huggingface', 'metadata_fields': ['path', 'n'], 'num_data_workers': 0,
                             'pin_memory': False, 'batch_size': 2})
    loader = with_meta.get_dataloader(tmp_path / 'ds', FakeEncoder())
    assert loader.dataset.metadata == [{'path': f'p{i}', 'n': i} for i in range(5)]
    assert loader.batch_size == 2


def test_sentence_splitter_choice_is_explicit():
    """Without nltk the regex stand-in is never silent: 'auto' warns, 'punkt' raises, 'regex' is the opt-in."""
    import importlib.util
    import warnings

    if importlib.util.find_spec('nltk') is not None:
        pytest.skip('nltk installed: the reference splitter is used')
    with pytest.warns(RuntimeWarning, match='nltk is not installed'):
        split_by_sentence_tokenizer('auto')
    with pytest.raises(ImportError, match='nltk is required'):
        split_by_sentence_tokenizer('punkt')
    with warnings.catch_warnings():
        warnings.simplefilter('error')
        assert ''.join(split_by_sentence_tokenizer('regex')('One. Two.')) == 'One. Two.'
    with pytest.raises(ValueError):
        split_by_sentence_tokenizer('other')


def test_ubinary_search_oracle_properties():
    """oracle/search.py's restatement of the reference's binary branch (quantize -> IndexBinaryFlat -> rescore):
    known-answer checks of every piece (no faiss / sentence_transformers here: parity unpinned, see its header)."""
    from oracle import search as osearch

    x = np.array([[0.5, -1.0, 0.0, 2.0, -0.1, 3.0, 1e-9, -7.0,     1.0, 1.0, 1.0, 1.0, -1.0, -1.0, -1.0, -1.0]],
                 dtype=np.float32)
    bits = osearch.quantize_ubinary(x)
    assert bits.dtype == np.uint8 and bits.tolist() == [[0b10010110, 0b11110000]]   # first dimension = MSB; 0 -> 0
    rng = np.random.default_rng(3)
    corpus = rng.standard_normal((300, 64)).astype(np.float32)
    corpus[17] = corpus[5]                       # identical rows: a Hamming tie, the smaller id first
    cb = osearch.quantize_ubinary(corpus)
    q = corpus[5:6] + 0.01 * rng.standard_normal((1, 64)).astype(np.float32)
    d, i = osearch.hamming_topk(osearch.quantize_ubinary(q), cb, 4)
    assert i[0, 0] == 5 and i[0, 1] == 17 and d[0, 0] == d[0, 1] and list(d[0]) == sorted(d[0])
    brute = (np.unpackbits(cb, axis=1) != np.unpackbits(osearch.quantize_ubinary(q), axis=1)).sum(1)
    assert sorted(brute)[:4] == list(d[0])
    s, idx = osearch.search_ubinary(q, cb, top_k=3, rescore_multiplier=4)
    cand = osearch.hamming_topk(osearch.quantize_ubinary(q), cb, 12)[1][0]
    resc = (np.unpackbits(cb[cand], axis=1) * q[0]).sum(1)
    assert np.allclose(s[0], np.sort(resc)[::-1][:3], rtol=1e-6) and set(idx[0]) <= set(cand)
    assert list(s[0]) == sorted(s[0], reverse=True)
    s2, idx2 = osearch.search_ubinary(q, cb[:2], top_k=3, rescore_multiplier=2)   # corpus smaller than k
    assert s2.shape == (1, 2)


def test_exact_index_config_accepts_the_reference_fields(tmp_path):
    from distllm_b200.rag.search import ExactIndexConfig

    cfg = ExactIndexConfig(name='faiss_index_v2', dataset_dir=tmp_path, faiss_index_path=tmp_path / 'x.index',
                           precision='ubinary', rescore_multiplier=4, num_quantization_workers=8)
    assert (cfg.precision, cfg.search_algorithm, cfg.rescore_multiplier) == ('ubinary', 'exact', 4)
    assert ExactIndexConfig().precision == 'float32' and ExactIndexConfig().rescore_multiplier == 2
    with pytest.raises(Exception):
        ExactIndexConfig(search_algorithm='hnsw')      # the approximate branch is not built
    with pytest.raises(Exception):
        ExactIndexConfig(precision='uint8')            # the reference itself only accepts float32 / ubinary


def test_nf4_roundtrip_structure():
    import torch

    """embed/encoders/nf4.py (restatement of bitsandbytes' NF4 + double quantisation, reference auto.py:44-56):
    every dequantised weight is (one of the 16 NF4 code values) x (its block's scale); errors stay within the
    spacing of the code; only nn.Linear weights of the blocks are touched."""
    from distllm_b200.embed.encoders.nf4 import NF4_CODE
    from distllm_b200.embed.encoders.nf4 import dynamic_map_8bit
    from distllm_b200.embed.encoders.nf4 import nf4_roundtrip
    from distllm_b200.embed.encoders.nf4 import quantize_state_dict_nf4

    code8 = dynamic_map_8bit()
    assert len(code8) == 256 and code8[-1] == 1.0 and (code8 == 0).sum() == 1 and bool((code8[1:] > code8[:-1]).all())
    assert len(NF4_CODE) == 16 and NF4_CODE[0] == -1.0 and NF4_CODE[7] == 0.0 and NF4_CODE[-1] == 1.0
    g = torch.Generator().manual_seed(0)
    w = torch.randn(96, 200, generator=g) * 0.02
    r = nf4_roundtrip(w, double_quant=False)
    blocks_w, blocks_r = w.flatten().view(-1, 64), r.flatten().view(-1, 64)
    absmax = blocks_w.abs().amax(1, keepdim=True)
    ratio = blocks_r / absmax
    code = torch.tensor(NF4_CODE)
    assert float((ratio[..., None] - code).abs().amin(-1).max()) < 1e-6      # code value x block absmax
    # nearest code value: the error is at most half the widest gap of the code, relative to the block scale
    widest = float((code[1:] - code[:-1]).max())
    assert float(((blocks_r - blocks_w).abs() / absmax).max()) <= widest / 2 + 1e-6
    # the block maximum itself is reproduced exactly (it maps to +-1)
    assert torch.allclose(blocks_r.abs().amax(1), absmax[:, 0], rtol=1e-6)
    # double quantisation perturbs the block scales by a fraction of a percent only
    r2 = nf4_roundtrip(w)
    assert 0 < float((r2 - r).norm() / r.norm()) < 0.02
    assert 0.05 < float((r2 - w).norm() / w.norm()) < 0.15
    sd = {'encoder.layer.0.attention.self.query.weight': w, 'encoder.layer.0.attention.self.query.bias': w[0],
          'embeddings.word_embeddings.weight': w, 'encoder.layer.0.output.LayerNorm.weight': w[0]}
    q = quantize_state_dict_nf4(sd)
    assert torch.equal(q['encoder.layer.0.attention.self.query.weight'], r2)
    assert all(q[k] is sd[k] for k in sd if 'query.weight' not in k)


@pytest.mark.parametrize('dtype', ['float16', 'bfloat16'])
def test_weight_lists_follow_the_abi_order_on_cpu(dtype):
    """embed/encoders/weights.py: every family's state dict -> ABI-ordered tensor list (count = b2e_num_weights,
    matrices in the build's 16-bit storage type, vectors fp32).  Runs on CPU tensors: no GPU needed."""
    import ctypes as C

    import torch
    from transformers import BertConfig
    from transformers import EsmConfig
    from transformers import MistralConfig
    from transformers import ModernBertConfig

    from distllm_b200 import _native
    from distllm_b200.embed.encoders import weights as W

    dt = getattr(torch, dtype)
    lib = _native.load(_native.storage_of(dt))
    cases = [
        (BertConfig(vocab_size=50, hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512,
                    max_position_embeddings=32), W.random_bert_state_dict, W.bert_desc, W.bert_weight_list),
        (EsmConfig(vocab_size=33, hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512,
                   max_position_embeddings=40, position_embedding_type='rotary', token_dropout=True, mask_token_id=32,
                   pad_token_id=1), W.random_esm_state_dict, W.esm_desc, W.esm_weight_list),
        (MistralConfig(vocab_size=50, hidden_size=256, num_hidden_layers=2, num_attention_heads=2,
                       num_key_value_heads=1, head_dim=128, intermediate_size=384, max_position_embeddings=64),
         W.random_mistral_state_dict, W.mistral_desc, W.mistral_weight_list),
        (ModernBertConfig(vocab_size=50, hidden_size=256, num_hidden_layers=4, num_attention_heads=4,
                          intermediate_size=200, max_position_embeddings=64, local_attention=16, pad_token_id=0,
                          bos_token_id=1, eos_token_id=2, cls_token_id=1, sep_token_id=2),
         W.random_modernbert_state_dict, W.modernbert_desc, W.modernbert_weight_list),
    ]
    for cfg, make, desc_fn, list_fn in cases:
        sd = make(cfg, seed=0, device='cpu')
        desc = desc_fn(cfg)
        tensors = list_fn(sd, cfg.num_hidden_layers, torch.device('cpu'), dt)
        assert len(tensors) == lib.b2e_num_weights(C.byref(desc)), type(cfg).__name__
        assert lib.b2e_check_model(C.byref(desc)) == 0, lib.b2e_last_error()
        assert all(t.is_contiguous() for t in tensors)
        assert {t.dtype for t in tensors if t.dim() == 2 and t.shape[0] != cfg.vocab_size
                and t.shape[0] != getattr(cfg, 'max_position_embeddings', -1)
                and t.shape[0] != getattr(cfg, 'type_vocab_size', -1)} == {dt}
        assert all(t.dtype == torch.float32 for t in tensors if t.dim() == 1)
    # ModernBERT: intermediate 200 is zero-padded to 256 (gelu(0) * 0 feeds zero columns of mlp.Wo)
    mb = cases[3]
    tensors = mb[3](mb[1](mb[0], seed=0, device='cpu'), 4, torch.device('cpu'), dt)
    assert tensors[5 + 6].shape == (512, 256) and tensors[5 + 7].shape == (256, 256)
    assert not tensors[5 + 7][:, 200:].any()


# ------------------------------------------------------------------- host feed (sentence token cache)
def _wordpiece_tokenizer(tmp_path, max_len=48):
    from transformers import BertTokenizerFast

    words = [f'w{i:03d}' for i in range(200)] + ['play', '##ing', '##ed', 'un', '##believ', '##able', 'the', 'a',
                                                 'cell', '##s', 'protein', '.', ',', '!', '?', '(', ')', '-', ';',
                                                 'é', 'e', '中', '文', 'x', '##x']
    (tmp_path / 'vocab.txt').write_text('\n'.join(['[PAD]', '[UNK]', '[CLS]', '[SEP]', '[MASK]', *words]) + '\n')
    tok = BertTokenizerFast(vocab=str(tmp_path / 'vocab.txt'), do_lower_case=True)
    tok.model_max_length = max_len
    assert tok.convert_tokens_to_ids('playing') == tok.unk_token_id and len(tok('playing')['input_ids']) == 4
    return tok


def _documents(n_docs=5, seed=3):
    import random

    rng = random.Random(seed)
    words = [f'w{i:03d}' for i in range(200)] + ['playing', 'played', 'unbelievable', 'The', 'cells', 'protein',
                                                 'Élan', '中文', 'xxxx', 'w001-w002', '(w003)', 'ét']
    seps = [' ', '  ', '\n', '\t', ' \n ', ' ']
    docs = []
    for d in range(n_docs):
        sentences = []
        for s in range(rng.randint(6, 30)):
            n = rng.choice([2, 5, 9, 14, 60]) if s % 7 else 80      # some sentences exceed max_length alone
            sentences.append('The ' + ' '.join(rng.choice(words) for _ in range(n)) + rng.choice(['.', '!', '?', '...']))
        text = ''
        for s in sentences:
            text += s + rng.choice(seps)
        docs.append(text)
    return docs


def test_sentence_token_cache_gives_the_tokenizers_own_ids(tmp_path):
    """jsonl_chunk buffers and semantic chunks assembled from once-tokenised sentences == tokenizer(text), id for
    id (padding, truncation, special tokens, type ids, attention mask), through the real dataset and collator."""
    import json
    from types import SimpleNamespace

    import torch

    from distllm_b200.embed.datasets.jsonl_chunk import JsonlChunkDataset
    from distllm_b200.embed.datasets.jsonl_chunk import JsonlChunkDatasetConfig
    from distllm_b200.embed.datasets.utils import DataCollator
    from distllm_b200.embed.datasets.utils import InMemoryDataset
    from distllm_b200.embed.datasets.utils import SentenceTokenCache

    tok = _wordpiece_tokenizer(tmp_path)
    assert SentenceTokenCache.supported(tok)
    path = tmp_path / 'docs.jsonl'
    with path.open('w') as f:
        for i, text in enumerate(_documents()):
            f.write(json.dumps({'text': text, 'path': f'doc{i}'}) + '\n')
    enc = SimpleNamespace(tokenizer=tok)
    loaders = {}
    for on in (True, False):
        cfg = JsonlChunkDatasetConfig(buffer_size=2, batch_size=7, num_data_workers=0, pin_memory=False,
                                      sentence_splitter='regex', sentence_token_cache=on, min_buffer_length=20)
        loaders[on] = JsonlChunkDataset(cfg).get_dataloader(path, enc)
    fast, plain = loaders[True], loaders[False]
    assert fast.dataset.token_cache is not None and plain.dataset.token_cache is None
    assert fast.dataset.data == plain.dataset.data and len(fast.dataset) > 40
    n_rows = 0
    for a, b in zip(fast, plain):
        assert set(a.keys()) == set(b.keys()) == {'input_ids', 'token_type_ids', 'attention_mask'}
        for key in a.keys():
            assert torch.equal(a[key], b[key]), key
        n_rows += a['input_ids'].shape[0]
        assert a['input_ids'].shape[1] <= tok.model_max_length
    assert n_rows == len(fast.dataset)
    # pass 2: chunks = runs of consecutive rows, joined from the sentences those rows are centred on
    ds = fast.dataset
    groups = [(0, 3), (3, 4), (4, 11), (11, len(ds))]
    texts = [''.join(m['sentence'] for m in ds.metadata[s:e]) for s, e in groups]
    parts = [tuple(ds.sentence_index[r] for r in range(s, e)) for s, e in groups]
    chunks = InMemoryDataset(texts, None, parts=parts, token_cache=ds.token_cache)
    got = DataCollator(tok, cache=chunks.token_cache)([chunks[i] for i in range(len(chunks))])
    want = DataCollator(tok)(texts)
    for key in want.keys():
        assert torch.equal(got[key], want[key]), key


def test_sentence_token_cache_refuses_what_it_cannot_prove(tmp_path):
    import torch

    from distllm_b200.embed.datasets.utils import DataCollator
    from distllm_b200.embed.datasets.utils import PieceText
    from distllm_b200.embed.datasets.utils import SentenceTokenCache

    tok = _wordpiece_tokenizer(tmp_path)
    # a joint inside a word: "un" + "believable" must NOT be assembled from two pieces
    sentences = ['w001 un', 'believable w002. ', 'w003 w004.']
    cache = SentenceTokenCache.build(tok, sentences)
    assert cache.row_ids((0, 1), 48) is None and cache.row_ids((1, 2), 48) is not None
    text = PieceText(''.join(sentences), (0, 1, 2))
    got = DataCollator(tok, cache=cache)([text])
    want = DataCollator(tok)([str(text)])
    assert torch.equal(got['input_ids'], want['input_ids'])     # fell back to tokenising the text
    # other tokenizer families are left alone
    llama, _ = _llama_like_tokenizer()
    assert not SentenceTokenCache.supported(llama) and SentenceTokenCache.build(llama, sentences) is None
    # DataLoader worker processes pickle the items
    import pickle

    back = pickle.loads(pickle.dumps(text))
    assert back == text and back.parts == (0, 1, 2)


def test_partition_host_threads_gives_each_local_rank_its_share(monkeypatch):
    import os

    import torch

    from distllm_b200.sharding import partition_host_threads

    before = torch.get_num_threads()
    try:
        monkeypatch.delenv('LOCAL_WORLD_SIZE', raising=False)
        monkeypatch.delenv('WORLD_SIZE', raising=False)
        monkeypatch.delenv('RAYON_NUM_THREADS', raising=False)
        monkeypatch.delenv('OMP_NUM_THREADS', raising=False)
        assert partition_host_threads() is None and 'RAYON_NUM_THREADS' not in os.environ
        monkeypatch.setenv('LOCAL_WORLD_SIZE', '4')
        cores = len(os.sched_getaffinity(0))
        share = partition_host_threads()
        assert share == max(1, cores // 4) and os.environ['RAYON_NUM_THREADS'] == str(share)
        assert torch.get_num_threads() == share
        # an explicit user setting wins
        monkeypatch.setenv('RAYON_NUM_THREADS', '3')
        partition_host_threads()
        assert os.environ['RAYON_NUM_THREADS'] == '3'
    finally:
        torch.set_num_threads(before)
