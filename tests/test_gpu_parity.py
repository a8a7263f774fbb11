"""-m gpu: the CUDA path against the CPU oracle / the reference's golden vectors.

north_star tolerance: pooled embeddings within 1e-3 cosine of the reference's fp32 CPU encoder
(the native path multiplies in fp16 with fp32 accumulation)."""

from __future__ import annotations

import numpy as np
import pytest
import torch
from transformers import BatchEncoding
from transformers import BertConfig

from distllm_b200 import _native as nv
from distllm_b200.embed.encoders.auto import AutoEncoder
from distllm_b200.embed.encoders.native import NativeBertEncoder
from distllm_b200.embed.encoders.weights import random_bert_state_dict
from distllm_b200.embed import get_embedder
from distllm_b200.embed import get_pooler
from oracle import bert as obert
from oracle import pooling as opool

from conftest import cosine_rows

pytestmark = pytest.mark.gpu
COS_TOL = 1e-3


class TokenBatches(torch.utils.data.Dataset):
    """Pre-tokenised batches behind the DataLoader interface the embedders consume."""

    def __init__(self, batches):
        self.batches = batches
        self.data = [f'row{i}' for i in range(sum(len(b['input_ids']) for b in batches))]
        self.metadata = None

    def __len__(self):
        return len(self.data)

    def __getitem__(self, i):
        return BatchEncoding(self.batches[i])


def loader_of(batches):
    ds = TokenBatches(batches)
    return torch.utils.data.DataLoader(ds, batch_size=None, sampler=range(len(batches)))


def golden_batches(golden):
    return [{k: torch.from_numpy(golden[f'batch{i}/{k}']) for k in ('input_ids', 'attention_mask', 'token_type_ids')}
            for i in range(int(golden['n_batches']))]


@pytest.fixture(scope='module')
def tiny_native(tiny_bert):
    cfg, sd = tiny_bert
    enc = NativeBertEncoder(cfg, sd)
    yield enc
    enc.close()


def test_hidden_state_matches_reference(tiny_native, bert_golden):
    b = golden_batches(bert_golden)[0]
    hidden = tiny_native.encode(b['input_ids'], b['attention_mask'], b['token_type_ids']).cpu().numpy()
    ref = bert_golden['batch0/hidden']
    assert hidden.shape == ref.shape
    cos = cosine_rows(hidden.reshape(-1, ref.shape[-1]), ref.reshape(-1, ref.shape[-1]))
    assert cos.min() > 1 - COS_TOL, cos.min()
    assert np.abs(hidden - ref).max() < 0.15  # post-LayerNorm values are O(1)


@pytest.mark.parametrize('kind,pooler,normalize', [('mean', 'mean', False), ('mean_normalized', 'mean', True),
                                                   ('last_token', 'last_token', False)])
@pytest.mark.parametrize('fused', [True, False])
def test_embedder_matches_reference_embeddings(tiny_native, bert_golden, kind, pooler, normalize, fused):
    """full_sequence embedder through the plugin API == the reference's compute_embeddings output,
    both via the fused encode_pooled path and via encode() + Pooler.pool()."""
    encoder = AutoEncoder.from_native(tiny_native)
    if not fused:
        class Unfused:  # hides encode_pooled so the embedder takes the generic path
            dtype, device, embedding_size = encoder.dtype, encoder.device, encoder.embedding_size
            tokenizer = None
            encode = staticmethod(encoder.encode)
        encoder = Unfused()
    embedder = get_embedder({'name': 'full_sequence', 'normalize_embeddings': normalize})
    result = embedder.embed(loader_of(golden_batches(bert_golden)), encoder, get_pooler({'name': pooler}))
    ref = bert_golden[f'pooled/{kind}']
    assert result.embeddings.shape == ref.shape and result.embeddings.dtype == np.float32
    cos = cosine_rows(result.embeddings, ref)
    assert cos.min() > 1 - COS_TOL, cos
    if normalize:
        np.testing.assert_allclose(np.linalg.norm(result.embeddings, axis=-1), 1.0, atol=1e-5)


def test_fused_pool_does_not_touch_the_mask_but_pooler_does(tiny_native, bert_golden):
    b = golden_batches(bert_golden)[1]
    mask = b['attention_mask'].cuda()
    keep = mask.clone()
    tiny_native.encode_pooled(b['input_ids'], mask, b['token_type_ids'], nv.POOL_MEAN_REF, False)
    assert torch.equal(mask, keep)
    hidden = tiny_native.encode(b['input_ids'], mask, b['token_type_ids'])
    get_pooler({'name': 'mean'}).pool(hidden, mask)
    ref_mask = b['attention_mask'].clone()
    opool.average_pool(torch.zeros(*ref_mask.shape, 1), ref_mask)
    assert torch.equal(mask.cpu(), ref_mask)


def test_embed_host_equals_device_path_bitwise(tiny_native, bert_golden):
    batches = golden_batches(bert_golden)
    s = max(b['input_ids'].shape[1] for b in batches)
    b0 = batches[2]  # one batch, fed as host tensors
    ids, mask, types = (b0[k].contiguous().pin_memory() for k in ('input_ids', 'attention_mask', 'token_type_ids'))
    host = tiny_native.embed_host(ids, mask, types, batch=ids.shape[0], pool_kind=nv.POOL_MEAN_REF, normalize=False)
    devp = tiny_native.encode_pooled(ids, mask, types, nv.POOL_MEAN_REF, False)
    torch.cuda.synchronize()
    assert torch.equal(host, devp.cpu())
    # several batches in one call: batch composition (hence the quirk) follows `batch`
    n = 2 * ids.shape[0]
    ids2, mask2, types2 = (torch.cat([t, t]).contiguous() for t in (ids, mask, types))
    host2 = tiny_native.embed_host(ids2, mask2, types2, batch=ids.shape[0], pool_kind=nv.POOL_MEAN_REF, normalize=False)
    assert torch.equal(host2[: n // 2], host) and torch.equal(host2[n // 2:], host)
    assert s >= 1


def test_embed_host_graph_replay_equals_eager(tiny_native, tiny_bert):
    """From the second full batch on b2e_embed_host replays a captured CUDA graph (two staging slots):
    every batch must equal the eager per-batch call bit for bit, including a ragged last batch, a
    second call with another shape (new graphs) and a return to the first shape (cached graphs)."""
    cfg, _ = tiny_bert
    g = torch.Generator().manual_seed(31)

    def make(n, s):
        ids = torch.randint(5, cfg.vocab_size, (n, s), generator=g)
        lens = torch.randint(1, s + 1, (n,), generator=g)
        mask = (torch.arange(s)[None] < lens[:, None]).long()
        return ids.pin_memory(), mask.pin_memory(), torch.zeros_like(ids).pin_memory()

    for n, s, batch, pool in [(27, 40, 4, nv.POOL_MEAN_REF), (16, 64, 8, nv.POOL_LAST_TOKEN),
                              (27, 40, 4, nv.POOL_MEAN_REF)]:
        ids, mask, types = make(n, s)
        host = tiny_native.embed_host(ids, mask, types, batch=batch, pool_kind=pool, normalize=True)
        for r0 in range(0, n, batch):
            sl = slice(r0, min(n, r0 + batch))
            ref = tiny_native.encode_pooled(ids[sl], mask[sl], types[sl], pool, True)
            assert torch.equal(host[sl], ref.cpu()), (n, s, batch, r0)


@pytest.fixture(scope='module')
def base_model():
    """BERT-base shape (S-PubMedBert-MS-MARCO: L=12, H=768, 12 heads, I=3072), seeded random weights."""
    cfg = BertConfig(vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                     intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2,
                     layer_norm_eps=1e-12, initializer_range=0.02)
    sd = random_bert_state_dict(cfg, seed=0, device='cpu')
    enc = NativeBertEncoder(cfg, sd)
    yield cfg, sd, enc
    enc.close()


def test_bert_base_pooled_vs_oracle(base_model):
    cfg, sd, enc = base_model
    g = torch.Generator().manual_seed(1)
    b, s = 6, 128
    ids = torch.randint(7, cfg.vocab_size, (b, s), generator=g)
    lens = torch.tensor([128, 128, 64, 9, 100, 31])
    mask = (torch.arange(s)[None] < lens[:, None]).long()
    hidden_ref = obert.bert_forward(sd, cfg, ids, mask)
    for kind, pool in ((nv.POOL_MEAN_REF, opool.average_pool), (nv.POOL_LAST_TOKEN, opool.last_token_pool)):
        ref = pool(hidden_ref, mask.clone()).numpy()
        got = enc.encode_pooled(ids, mask, None, kind, False).cpu().numpy()
        cos = cosine_rows(got, ref)
        assert cos.min() > 1 - COS_TOL, (kind, cos)
    hidden = enc.encode(ids, mask).cpu().numpy()
    valid = mask.bool().numpy()
    cos_tok = cosine_rows(hidden[valid], hidden_ref.numpy()[valid])
    assert cos_tok.min() > 1 - COS_TOL, cos_tok.min()


def test_full_size_properties(base_model):
    """BASELINE config C2 sizes (S=512, batch cut to 64 rows): properties that need no oracle."""
    cfg, _, enc = base_model
    g = torch.Generator().manual_seed(2)
    b, s = 64, 512
    ids = torch.randint(7, cfg.vocab_size, (b, s), generator=g).cuda()
    lens = torch.randint(64, 513, (b,), generator=g)
    lens[0] = 512
    mask = (torch.arange(s)[None] < lens[:, None]).long().cuda()
    a = enc.encode_pooled(ids, mask, None, nv.POOL_MEAN_PER_ROW, True)
    # determinism
    assert torch.equal(a, enc.encode_pooled(ids, mask, None, nv.POOL_MEAN_PER_ROW, True))
    # unit norm
    torch.testing.assert_close(a.norm(dim=-1), torch.ones(b, device=a.device), rtol=0, atol=1e-5)
    # rows are independent: permuting the batch permutes the result bit for bit
    perm = torch.randperm(b, generator=g).cuda()
    assert torch.equal(enc.encode_pooled(ids[perm], mask[perm], None, nv.POOL_MEAN_PER_ROW, True), a[perm])
    # padding invariance: a sequence alone at its own length == inside the padded batch
    i = 5
    n = int(lens[i])
    alone = enc.encode_pooled(ids[i:i + 1, :n].contiguous(), mask[i:i + 1, :n].contiguous(), None,
                              nv.POOL_MEAN_PER_ROW, True)
    assert torch.nn.functional.cosine_similarity(alone, a[i:i + 1]).item() > 1 - 1e-5
    # reference-quirk pooling on an unpadded batch == "drop first and last token" per row
    full = torch.ones(8, s, dtype=torch.int64, device='cuda')
    q = enc.encode_pooled(ids[:8], full, None, nv.POOL_MEAN_REF, False)
    p = enc.encode_pooled(ids[:8], full, None, nv.POOL_MEAN_PER_ROW, False)
    assert torch.equal(q, p)


def test_semantic_chunk_embedder_end_to_end(tmp_path, tiny_bert, tiny_native):
    """jsonl_chunk dataset -> semantic_chunk embedder through the plugin API with a real tokenizer.
    The split is checked against the oracle on the SAME pass-1 embeddings (exact), the final chunk
    embeddings against the oracle forward (cosine)."""
    import json

    from transformers import BertTokenizerFast

    from distllm_b200.embed import get_dataset
    from distllm_b200.embed.embedders import semantic_chunk as sc
    from distllm_b200.embed.embedders.full_sequence import compute_embeddings_device
    from oracle import semantic as osem
    from oracle.make_golden import TINY

    cfg, sd = tiny_bert
    words = [f'w{i:03d}' for i in range(TINY['vocab_size'] - 5)]
    (tmp_path / 'vocab.txt').write_text('\n'.join(['[PAD]', '[UNK]', '[CLS]', '[SEP]', '[MASK]', *words]) + '\n')
    tok = BertTokenizerFast(vocab=str(tmp_path / 'vocab.txt'), do_lower_case=False)
    tok.model_max_length = cfg.max_position_embeddings
    rng = np.random.default_rng(0)
    docs = []
    for d in range(3):
        sents = [('S' + ' '.join(rng.choice(words, size=rng.integers(5, 9))) + '. ') for _ in range(12 + d)]
        docs.append({'text': ''.join(sents), 'path': f'doc{d}'})
    f = tmp_path / 'docs.jsonl'
    f.write_text('\n'.join(json.dumps(d) for d in docs))

    encoder = AutoEncoder.from_native(tiny_native, tokenizer=tok)
    dataset = get_dataset({'name': 'jsonl_chunk', 'buffer_size': 1, 'min_buffer_length': 20, 'batch_size': 5,
                           'num_data_workers': 0, 'pin_memory': False})
    pooler = get_pooler({'name': 'mean'})
    loader = dataset.get_dataloader(f, encoder)
    n_buffers = len(loader.dataset)
    assert n_buffers == sum(12 + d for d in range(3))
    pass1 = compute_embeddings_device(loader, encoder, pooler, progress=False)
    ranges = sc.document_ranges(loader.dataset.metadata)
    want_groups = osem.split_rows(pass1.cpu().numpy(), ranges, 80)

    embedder = get_embedder({'name': 'semantic_chunk', 'breakpoint_percentile_threshold': 80,
                             'chunk_batch_size': 4, 'min_chunk_length': 10})
    loader = dataset.get_dataloader(f, encoder)  # fresh metadata ('sentence' is popped in place)
    sentences = [m['sentence'] for m in loader.dataset.metadata]
    result = embedder.embed(loader, encoder, pooler)
    want_texts = [''.join(sentences[s:e]) for s, e in want_groups]
    assert result.text == want_texts
    assert all('sentence' not in m for m in result.metadata)
    assert result.embeddings.shape == (len(want_texts), cfg.hidden_size)

    # final embeddings vs the oracle forward on the same chunk texts, same batching (4 per batch)
    ref = []
    for i in range(0, len(want_texts), 4):
        enc_b = tok(want_texts[i:i + 4], padding=True, truncation=True, return_tensors='pt')
        hidden = obert.bert_forward(sd, cfg, enc_b['input_ids'], enc_b['attention_mask'], enc_b['token_type_ids'])
        ref.append(opool.average_pool(hidden, enc_b['attention_mask'].clone()))
    cos = cosine_rows(result.embeddings, torch.cat(ref).numpy())
    assert cos.min() > 1 - COS_TOL, cos


# ---------------------------------------------------------------------------------- ESM-2
def test_esm2_matches_reference_vectors(tiny_esm, esm_golden):
    """esm2 encoder through the plugin API vs the reference's Esm2Encoder outputs (rotary positions,
    token dropout rows, a row truncated to max_position_embeddings, S up to 160 > 128)."""
    from distllm_b200.embed.encoders.esm2 import Esm2Encoder
    from distllm_b200.embed.encoders.native import NativeEsm2Encoder

    cfg, sd = tiny_esm
    native = NativeEsm2Encoder(cfg, sd)
    try:
        encoder = Esm2Encoder.from_native(native)
        batches = [{k: torch.from_numpy(esm_golden[f'batch{i}/{k}']) for k in ('input_ids', 'attention_mask')}
                   for i in range(int(esm_golden['n_batches']))]
        hidden = encoder.encode(BatchEncoding(batches[0])).cpu().numpy()
        ref = esm_golden['batch0/hidden']
        valid = batches[0]['attention_mask'].bool().numpy()
        cos = cosine_rows(hidden[valid], ref[valid])
        assert cos.min() > 1 - COS_TOL, cos.min()
        for fused in (True, False):
            enc = encoder
            if not fused:
                class Unfused:
                    dtype, device, embedding_size = encoder.dtype, encoder.device, encoder.embedding_size
                    tokenizer = None
                    encode = staticmethod(encoder.encode)
                enc = Unfused()
            result = get_embedder({'name': 'full_sequence'}).embed(loader_of(batches), enc, get_pooler({'name': 'mean'}))
            cos = cosine_rows(result.embeddings, esm_golden['pooled/mean'])
            assert cos.min() > 1 - COS_TOL, (fused, cos)
    finally:
        native.close()


def test_esm2_650m_width_vs_oracle():
    """ESM2-650M layer shape (H=1280, 20 heads, I=5120) with 3 layers, S=300 ragged, vs the CPU oracle."""
    from transformers import EsmConfig

    from distllm_b200.embed.encoders.native import NativeEsm2Encoder
    from distllm_b200.embed.encoders.weights import random_esm_state_dict
    from oracle import esm as oesm

    cfg = EsmConfig(vocab_size=33, hidden_size=1280, num_hidden_layers=3, num_attention_heads=20,
                    intermediate_size=5120, max_position_embeddings=1026, position_embedding_type='rotary',
                    token_dropout=True, mask_token_id=32, pad_token_id=1, layer_norm_eps=1e-5,
                    emb_layer_norm_before=False, initializer_range=0.02)
    sd = random_esm_state_dict(cfg, seed=3, device='cpu')
    g = torch.Generator().manual_seed(4)
    b, s = 3, 300
    ids = torch.randint(4, 24, (b, s), generator=g)
    lens = torch.tensor([300, 41, 177])
    mask = (torch.arange(s)[None] < lens[:, None]).long()
    ids = ids.masked_fill(mask == 0, 1)
    ids[:, 0] = 0
    ref_hidden = oesm.esm_forward(sd, cfg, ids, mask)
    ref = opool.average_pool(ref_hidden, mask.clone()).numpy()
    native = NativeEsm2Encoder(cfg, sd)
    try:
        got = native.encode_pooled(ids, mask, None, nv.POOL_MEAN_REF, False).cpu().numpy()
        cos = cosine_rows(got, ref)
        assert cos.min() > 1 - COS_TOL, cos
        hidden = native.encode(ids, mask).cpu().numpy()
        valid = mask.bool().numpy()
        assert cosine_rows(hidden[valid], ref_hidden.numpy()[valid]).min() > 1 - COS_TOL
    finally:
        native.close()


# ---------------------------------------------------------------------------------- Mistral family
@pytest.mark.parametrize('variant', ['full', 'window'])
@pytest.mark.parametrize('side', ['right', 'left'])
def test_mistral_matches_reference_vectors(mistral_golden, variant, side):
    """`auto` encoder on a Mistral checkpoint through the plugin API vs the reference's AutoEncoder
    (HF MistralModel) outputs: grouped-query causal attention (head_dim 128), rotary, RMSNorm, SwiGLU;
    right- and left-padded batches, without / with a sliding window, a row truncated at 320 tokens."""
    from conftest import tiny_mistral_variant

    from distllm_b200.embed.encoders.native import NativeMistralEncoder

    cfg, sd = tiny_mistral_variant(variant)
    key = f'{variant}/{side}'
    native = NativeMistralEncoder(cfg, sd)
    try:
        encoder = AutoEncoder.from_native(native)
        batches = [{k: torch.from_numpy(mistral_golden[f'{key}/batch{i}/{k}'])
                    for k in ('input_ids', 'attention_mask')} for i in range(int(mistral_golden['n_batches']))]
        if f'{key}/batch1/hidden' in mistral_golden.files:
            hidden = encoder.encode(BatchEncoding(batches[1])).cpu().numpy()
            ref = mistral_golden[f'{key}/batch1/hidden']
            valid = batches[1]['attention_mask'].bool().numpy()
            assert np.isfinite(hidden).all()
            cos = cosine_rows(hidden[valid], ref[valid])
            assert cos.min() > 1 - COS_TOL, cos.min()
        for fused in (True, False):
            enc = encoder
            if not fused:
                class Unfused:
                    dtype, device, embedding_size = encoder.dtype, encoder.device, encoder.embedding_size
                    tokenizer = None
                    encode = staticmethod(encoder.encode)
                enc = Unfused()
            result = get_embedder({'name': 'full_sequence'}).embed(
                loader_of(batches), enc, get_pooler({'name': 'last_token'}))
            cos = cosine_rows(result.embeddings, mistral_golden[f'{key}/pooled/last_token'])
            assert cos.min() > 1 - COS_TOL, (fused, cos)
            if side == 'right':
                result = get_embedder({'name': 'full_sequence', 'normalize_embeddings': True}).embed(
                    loader_of(batches), enc, get_pooler({'name': 'mean'}))
                ref = mistral_golden[f'{key}/pooled/mean_normalized']
                # the 2-token row loses both tokens to the mean pooler (mean.py:35-36): exact zeros
                empty = np.linalg.norm(ref, axis=-1) == 0
                assert empty.sum() == 1 and not result.embeddings[empty].any()
                cos = cosine_rows(result.embeddings[~empty], ref[~empty])
                assert cos.min() > 1 - COS_TOL, (fused, cos)
    finally:
        native.close()


def test_mistral_7b_width_vs_oracle():
    """Mistral-7B layer shape (H=4096, 32 query / 8 kv heads x 128, I=14336) with 2 layers, S=400
    right-padded, vs the CPU oracle; last-token pooling as in BASELINE config C3."""
    from transformers import MistralConfig

    from distllm_b200.embed.encoders.native import NativeMistralEncoder
    from distllm_b200.embed.encoders.weights import random_mistral_state_dict
    from oracle import mistral as omis

    cfg = MistralConfig(vocab_size=2000, hidden_size=4096, num_hidden_layers=2, num_attention_heads=32,
                        num_key_value_heads=8, head_dim=128, intermediate_size=14336,
                        max_position_embeddings=4096, rms_norm_eps=1e-5, sliding_window=4096,
                        initializer_range=0.02)
    sd = random_mistral_state_dict(cfg, seed=5, device='cpu')
    g = torch.Generator().manual_seed(6)
    b, s = 3, 400
    ids = torch.randint(3, 2000, (b, s), generator=g)
    lens = torch.tensor([400, 57, 263])
    mask = (torch.arange(s)[None] < lens[:, None]).long()
    ref_hidden = omis.mistral_forward(sd, cfg, ids, mask)
    ref = opool.last_token_pool(ref_hidden, mask).numpy()
    native = NativeMistralEncoder(cfg, sd)
    try:
        got = native.encode_pooled(ids, mask, None, nv.POOL_LAST_TOKEN, False).cpu().numpy()
        cos = cosine_rows(got, ref)
        assert cos.min() > 1 - COS_TOL, cos
        hidden = native.encode(ids, mask).cpu().numpy()
        valid = mask.bool().numpy()
        assert cosine_rows(hidden[valid], ref_hidden.numpy()[valid]).min() > 1 - COS_TOL
    finally:
        native.close()


def test_mistral_long_sequence_properties():
    """S = 4096 (BASELINE C3 length), 1 layer at 7B width: determinism, and the last-token embedding of a
    right-padded row equals the embedding of the same row without its padding (causal attention never
    looks right), bit for bit."""
    from transformers import MistralConfig

    from distllm_b200.embed.encoders.native import NativeMistralEncoder
    from distllm_b200.embed.encoders.weights import random_mistral_state_dict

    cfg = MistralConfig(vocab_size=1000, hidden_size=4096, num_hidden_layers=1, num_attention_heads=32,
                        num_key_value_heads=8, head_dim=128, intermediate_size=14336,
                        max_position_embeddings=4096, rms_norm_eps=1e-5, sliding_window=4096,
                        initializer_range=0.02)
    dev = torch.device('cuda:0')
    sd = random_mistral_state_dict(cfg, seed=7, device=dev, dtype=torch.float16)
    native = NativeMistralEncoder(cfg, sd)
    try:
        g = torch.Generator().manual_seed(8)
        ids = torch.randint(3, 1000, (2, 4096), generator=g)
        mask = torch.ones(2, 4096, dtype=torch.int64)
        mask[1, 3000:] = 0
        a = native.encode_pooled(ids, mask, None, nv.POOL_LAST_TOKEN, True)
        b = native.encode_pooled(ids, mask, None, nv.POOL_LAST_TOKEN, True)
        assert torch.equal(a, b)
        assert torch.isfinite(a).all()
        np.testing.assert_allclose(a.norm(dim=-1).cpu().numpy(), 1.0, rtol=1e-5)
        short = native.encode_pooled(ids[1:, :3000].contiguous(), mask[1:, :3000].contiguous(), None,
                                     nv.POOL_LAST_TOKEN, True)
        cos = cosine_rows(short.cpu().numpy(), a[1:].cpu().numpy())
        assert cos.min() > 1 - 1e-5, cos
    finally:
        native.close()


# ---------------------------------------------------------------------------------- retrieval query path
def test_retriever_query_path_end_to_end(tmp_path, tiny_bert, tiny_native):
    """Retriever.search through the plugin surface (tokenise -> native encode -> mean pool -> normalise ->
    exact top-k on the device) vs the oracle forward + numpy search on the same corpus."""
    from transformers import BertTokenizerFast

    from distllm_b200.rag import ExactIndex
    from distllm_b200.rag import Retriever
    from oracle import search as osearch
    from oracle.make_golden import TINY

    cfg, sd = tiny_bert
    words = [f'w{i:03d}' for i in range(TINY['vocab_size'] - 5)]
    (tmp_path / 'vocab.txt').write_text('\n'.join(['[PAD]', '[UNK]', '[CLS]', '[SEP]', '[MASK]', *words]) + '\n')
    tok = BertTokenizerFast(vocab=str(tmp_path / 'vocab.txt'), do_lower_case=False)
    tok.model_max_length = cfg.max_position_embeddings
    encoder = AutoEncoder.from_native(tiny_native, tokenizer=tok)
    pooler = get_pooler({'name': 'mean'})
    rng = np.random.default_rng(5)
    corpus = rng.standard_normal((3000, cfg.hidden_size)).astype(np.float32)
    corpus /= np.linalg.norm(corpus, axis=1, keepdims=True)
    retriever = Retriever(encoder, pooler, ExactIndex(corpus), batch_size=4)
    queries = [' '.join(rng.choice(words, size=n)) for n in (5, 30, 12, 3, 50, 8, 20)]
    results, q_emb = retriever.search(queries, top_k=7)
    assert q_emb.shape == (7, cfg.hidden_size) and q_emb.dtype == np.float32
    np.testing.assert_allclose(np.linalg.norm(q_emb, axis=1), 1.0, rtol=1e-5)
    # the query embeddings against the oracle forward, same batching (sorted by length, batches of 4)
    order = sorted(range(len(queries)), key=lambda i: len(queries[i]))
    ref_rows = {}
    for b0 in range(0, len(order), 4):
        ids = [order[i] for i in range(b0, min(b0 + 4, len(order)))]
        enc = tok([queries[i] for i in ids], padding=True, truncation=True, return_tensors='pt')
        hidden = obert.bert_forward(sd, cfg, enc['input_ids'], enc['attention_mask'], enc['token_type_ids'])
        pooled = opool.average_pool(hidden, enc['attention_mask'].clone()).numpy()
        for i, row in zip(ids, osearch.normalize_l2(pooled)):
            ref_rows[i] = row
    ref_q = np.stack([ref_rows[i] for i in range(len(queries))])
    assert cosine_rows(q_emb, ref_q).min() > 1 - COS_TOL
    # the search itself: exact on the embeddings the retriever produced
    ref_s, ref_i = osearch.topk_inner_product(q_emb, corpus, 7)
    np.testing.assert_allclose(np.array(results.total_scores), ref_s, rtol=0, atol=2e-5)
    assert np.array(results.total_indices).tolist() == ref_i.tolist()
    # a precomputed embedding and a score threshold
    res2, _ = retriever.search(query_embedding=q_emb[:2], top_k=7, score_threshold=float(ref_s[0, 3]))
    assert res2.total_indices[0] == ref_i[0, :4].tolist()
    with pytest.raises(ValueError, match='at least one of'):
        retriever.search()


def test_esm2_full_length_properties():
    """BASELINE config C5 length (1024 residues -> S = 1026, not a multiple of the 128-row query tile nor of
    the 64-key chunk) at the 650M layer shape, 2 layers: determinism, unit norm, and the embedding of a
    right-padded sequence equals the embedding of the same sequence fed without its padding (the per-row
    mean pooler drops the padding; the encoder must not let padded keys or rows leak)."""
    from transformers import EsmConfig

    from distllm_b200.embed.encoders.native import NativeEsm2Encoder
    from distllm_b200.embed.encoders.weights import random_esm_state_dict

    cfg = EsmConfig(vocab_size=33, hidden_size=1280, num_hidden_layers=2, num_attention_heads=20,
                    intermediate_size=5120, max_position_embeddings=1026, position_embedding_type='rotary',
                    token_dropout=True, mask_token_id=32, pad_token_id=1, layer_norm_eps=1e-5,
                    emb_layer_norm_before=False, initializer_range=0.02)
    dev = torch.device('cuda:0')
    native = NativeEsm2Encoder(cfg, random_esm_state_dict(cfg, seed=11, device=dev))
    try:
        g = torch.Generator().manual_seed(12)
        s = 1026
        ids = torch.randint(4, 24, (3, s), generator=g)
        ids[:, 0] = 0
        lens = torch.tensor([1026, 700, 65])
        mask = (torch.arange(s)[None] < lens[:, None]).long()
        ids = ids.masked_fill(mask == 0, 1)
        a = native.encode_pooled(ids, mask, None, nv.POOL_MEAN_PER_ROW, True)
        b = native.encode_pooled(ids, mask, None, nv.POOL_MEAN_PER_ROW, True)
        assert torch.equal(a, b) and torch.isfinite(a).all()
        np.testing.assert_allclose(a.norm(dim=-1).cpu().numpy(), 1.0, rtol=1e-5)
        for row, n in ((1, 700), (2, 65)):
            alone = native.encode_pooled(ids[row:row + 1, :n].contiguous(), mask[row:row + 1, :n].contiguous(),
                                         None, nv.POOL_MEAN_PER_ROW, True)
            cos = cosine_rows(alone.cpu().numpy(), a[row:row + 1].cpu().numpy())
            assert cos.min() > 1 - 1e-5, (row, cos)
    finally:
        native.close()


# ---------------------------------------------------------------------------------- ModernBERT
@pytest.mark.parametrize('storage', ['bf16', 'f16'])
def test_modernbert_matches_reference_vectors(tiny_modernbert, modernbert_golden, storage):
    """`auto` encoder on a ModernBERT checkpoint through the plugin API vs the reference's AutoEncoder (HF
    ModernBertModel) outputs: rotary per layer type, alternating full / sliding-window attention (|i - j| <= 16
    here), GeGLU, pre-LN with an un-normed first layer; a 302-token row spans three query tiles."""
    from distllm_b200.embed.encoders.native import NativeModernBertEncoder

    cfg, sd = tiny_modernbert
    g = modernbert_golden
    native = NativeModernBertEncoder(cfg, sd, storage=storage)
    try:
        encoder = AutoEncoder.from_native(native)
        batches = [{k: torch.from_numpy(g[f'batch{i}/{k}']) for k in ('input_ids', 'attention_mask')}
                   for i in range(int(g['n_batches']))]
        hidden = encoder.encode(BatchEncoding(batches[1])).cpu().numpy()
        valid = batches[1]['attention_mask'].bool().numpy()
        assert np.isfinite(hidden).all()
        cos = cosine_rows(hidden[valid], g['batch1/hidden'][valid])
        assert cos.min() > 1 - COS_TOL, cos.min()
        for fused in (True, False):
            enc = encoder
            if not fused:
                class Unfused:
                    dtype, device, embedding_size = encoder.dtype, encoder.device, encoder.embedding_size
                    tokenizer = None
                    encode = staticmethod(encoder.encode)
                enc = Unfused()
            result = get_embedder({'name': 'full_sequence'}).embed(
                loader_of(batches), enc, get_pooler({'name': 'last_token'}))
            cos = cosine_rows(result.embeddings, g['pooled/last_token'])
            assert cos.min() > 1 - COS_TOL, (fused, cos)
            result = get_embedder({'name': 'full_sequence', 'normalize_embeddings': True}).embed(
                loader_of(batches), enc, get_pooler({'name': 'mean'}))
            ref = g['pooled/mean_normalized']
            empty = np.linalg.norm(ref, axis=-1) == 0
            assert not result.embeddings[empty].any()
            cos = cosine_rows(result.embeddings[~empty], ref[~empty])
            assert cos.min() > 1 - COS_TOL, (fused, cos)
    finally:
        native.close()


# ---------------------------------------------------------------------------------- padding-free layout
@pytest.mark.parametrize('family', ['bert', 'esm', 'modernbert', 'mistral', 'mistral-window'])
def test_packed_token_layout_equals_padded(family, tiny_bert, tiny_esm, tiny_modernbert):
    """Pooled forward passes run on the attended tokens only (csrc/pack.cuh).  Same batch, packing on vs off
    (b2e_debug_set_packing): right-padded ragged batches must give the same embeddings (padded positions can
    not influence a pooled row), batches with left padding / holes / empty rows fall back to the padded layout
    by themselves."""
    import ctypes

    from distllm_b200.embed.encoders.native import NativeEsm2Encoder
    from distllm_b200.embed.encoders.native import NativeModernBertEncoder

    from distllm_b200.embed.encoders.native import NativeMistralEncoder
    from conftest import tiny_mistral_variant

    if family.startswith('mistral'):
        cfg, sd = tiny_mistral_variant('window' if family.endswith('window') else 'full')
        cls = NativeMistralEncoder
    else:
        cfg, sd = {'bert': tiny_bert, 'esm': tiny_esm, 'modernbert': tiny_modernbert}[family]
        cls = {'bert': NativeBertEncoder, 'esm': NativeEsm2Encoder, 'modernbert': NativeModernBertEncoder}[family]
    enc = cls(cfg, sd)
    lib = enc._lib
    lib.b2e_debug_set_packing.argtypes = [ctypes.c_int]
    g = torch.Generator().manual_seed(5)
    s_max = min(cfg.max_position_embeddings, 300)
    try:
        for b, s, lens in [(7, 50, [50, 3, 17, 50, 1, 33, 2]), (5, s_max, [s_max, 129, 128, 64, 7]),
                           (3, 40, [40, 40, 40])]:
            lens = [min(n, s) for n in lens]
            ids = torch.randint(4, min(cfg.vocab_size, 24) if family == 'esm' else cfg.vocab_size - 1, (b, s),
                                generator=g)
            mask = (torch.arange(s)[None] < torch.tensor(lens)[:, None]).long()
            for kind in (nv.POOL_MEAN_REF, nv.POOL_MEAN_PER_ROW, nv.POOL_LAST_TOKEN):
                lib.b2e_debug_set_packing(1)
                packed = enc.encode_pooled(ids, mask, None, kind, True).clone()
                lib.b2e_debug_set_packing(0)
                padded = enc.encode_pooled(ids, mask, None, kind, True).clone()
                assert torch.isfinite(packed).all()
                live = padded.norm(dim=-1) > 0
                assert torch.equal(live, packed.norm(dim=-1) > 0)
                cos = torch.nn.functional.cosine_similarity(packed[live], padded[live])
                assert cos.min().item() > 1 - 1e-6, (family, b, s, kind, cos)
        # masks the packer must refuse (identity layout): results are then bit-identical by construction
        s = 40
        ids = torch.randint(4, min(cfg.vocab_size, 24) if family == 'esm' else cfg.vocab_size - 1, (4, s), generator=g)
        mask = torch.ones(4, s, dtype=torch.int64)
        mask[0, :10] = 0          # left padding
        mask[1, 5:9] = 0          # a hole
        mask[2, 20:] = 0
        lib.b2e_debug_set_packing(1)
        a = enc.encode_pooled(ids, mask, None, nv.POOL_MEAN_PER_ROW, False).clone()
        lib.b2e_debug_set_packing(0)
        b_ = enc.encode_pooled(ids, mask, None, nv.POOL_MEAN_PER_ROW, False).clone()
        assert torch.equal(a, b_)
    finally:
        lib.b2e_debug_set_packing(1)
        enc.close()
