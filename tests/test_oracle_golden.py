"""Pin the CPU oracle: against vectors produced by the unmodified reference (tests/golden, made by
oracle/make_golden.py) and against HuggingFace BertModel, which holds the reference's arithmetic."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import bert as obert
from oracle import pooling as opool
from oracle import semantic as osem
from oracle.make_golden import weights_digest

POOL_CASES = ['ragged', 'full', 'single', 'left_padded_like']


@pytest.mark.parametrize('case', POOL_CASES)
def test_average_pool_matches_reference(pool_golden, case):
    emb = torch.from_numpy(pool_golden[f'{case}/emb'])
    mask = torch.from_numpy(pool_golden[f'{case}/mask']).clone()
    got = opool.average_pool(emb, mask)
    np.testing.assert_allclose(got.numpy(), pool_golden[f'{case}/mean'], rtol=1e-5, atol=1e-6)
    # the in-place edit of the caller's mask is part of the contract
    np.testing.assert_array_equal(mask.numpy(), pool_golden[f'{case}/mask_after'])


@pytest.mark.parametrize('case', ['full', 'single', 'left_padded_like', 'leftpad'])
def test_last_token_pool_matches_reference(pool_golden, case):
    emb = torch.from_numpy(pool_golden[f'{case}/emb'])
    mask = torch.from_numpy(pool_golden[f'{case}/mask'])
    got = opool.last_token_pool(emb, mask)
    np.testing.assert_array_equal(got.numpy(), pool_golden[f'{case}/last'])


def test_mean_quirk_is_cross_row(pool_golden):
    """mean.py:36 clears column len_j-1 of EVERY row: row 0 (len 20) loses columns 4, 0, 1, 10, 18
    and, through the zero-length row's index -1, its own last column."""
    after = pool_golden['ragged/mask_after']
    assert after[0].sum() == 20 - len({0, 19, 4, 1, 10, 18})


def test_semantic_distances_and_groups_match_reference(semantic_golden):
    emb = semantic_golden['emb']
    for k, (lo, hi) in enumerate(semantic_golden['doc_ranges']):
        d = osem.calculate_distances_between_buffer(emb[lo:hi])
        assert d.dtype == np.float64
        np.testing.assert_allclose(d, semantic_golden[f'dist/{k}'], rtol=0, atol=2e-7)
        for pct in (50, 90, 95):
            groups = osem.build_chunks(semantic_golden[f'dist/{k}'], pct)
            assert [tuple(g) for g in semantic_golden[f'groups/{k}/{pct}']] == groups


def test_semantic_edge_documents(semantic_golden):
    # 1-row document: no distances -> the empty group (0, 0); 2-row document: one distance,
    # never strictly above its own percentile -> a single group
    assert osem.build_chunks(np.zeros(0), 90) == [(0, 0)]
    assert [tuple(g) for g in semantic_golden['groups/1/90']] == [(0, 0)]
    assert [tuple(g) for g in semantic_golden['groups/2/90']] == [(0, 2)]


def test_tiny_weights_are_reproducible(bert_golden, tiny_bert):
    _, sd = tiny_bert
    assert weights_digest(sd) == str(bert_golden['weights_sha256'])


def _batches(golden):
    for i in range(int(golden['n_batches'])):
        yield {k: torch.from_numpy(golden[f'batch{i}/{k}'])
               for k in ('input_ids', 'attention_mask', 'token_type_ids')}


def test_oracle_forward_matches_reference_hidden(bert_golden, tiny_bert):
    cfg, sd = tiny_bert
    batch = next(_batches(bert_golden))
    hidden = obert.bert_forward(sd, cfg, batch['input_ids'], batch['attention_mask'],
                                batch['token_type_ids'])
    np.testing.assert_allclose(hidden.numpy(), bert_golden['batch0/hidden'], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize('kind', ['mean', 'mean_normalized', 'last_token'])
def test_oracle_loop_matches_reference_embeddings(bert_golden, tiny_bert, kind):
    """oracle forward + oracle pooler + oracle loop == the reference's compute_embeddings output."""
    cfg, sd = tiny_bert

    def encode(batch):
        return obert.bert_forward(sd, cfg, batch['input_ids'], batch['attention_mask'],
                                  batch['token_type_ids'])

    pool = opool.last_token_pool if kind == 'last_token' else opool.average_pool
    got = opool.compute_embeddings(_batches(bert_golden), encode, pool,
                                   do_normalize=(kind == 'mean_normalized'))
    ref = bert_golden[f'pooled/{kind}']
    assert got.shape == ref.shape == (int(bert_golden['n_texts']), cfg.hidden_size)
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=2e-5)


def test_oracle_forward_matches_hf_bertmodel(tiny_bert):
    """Independent of the fixtures: the restatement against transformers' BertModel, ragged batch."""
    from transformers import BertModel

    cfg, sd = tiny_bert
    model = BertModel(cfg)
    model.load_state_dict(sd, strict=False)
    model.eval()
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(5, cfg.vocab_size, (5, 33), generator=g)
    lens = torch.tensor([33, 1, 17, 2, 30])
    mask = (torch.arange(33)[None] < lens[:, None]).long()
    types = torch.randint(0, 2, (5, 33), generator=g)
    with torch.no_grad():
        ref = model(input_ids=ids, attention_mask=mask, token_type_ids=types,
                    output_hidden_states=True).hidden_states
    got = obert.bert_forward(sd, cfg, ids, mask, types, return_all=True)
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-4, atol=2e-5)


# ---------------------------------------------------------------------------------- ESM-2
def _esm_batches(golden):
    for i in range(int(golden['n_batches'])):
        yield {k: torch.from_numpy(golden[f'batch{i}/{k}']) for k in ('input_ids', 'attention_mask')}


def test_esm_weights_are_reproducible(esm_golden, tiny_esm):
    _, sd = tiny_esm
    assert weights_digest(sd) == str(esm_golden['weights_sha256'])


def test_esm_oracle_matches_reference_hidden_and_embeddings(esm_golden, tiny_esm):
    """oracle ESM-2 forward (+ reference-semantics mean pool) == distllm's Esm2Encoder outputs,
    including the rows that contain <mask> tokens (token dropout) and the truncated 160-token row."""
    from oracle import esm as oesm

    cfg, sd = tiny_esm
    batches = list(_esm_batches(esm_golden))
    hidden = oesm.esm_forward(sd, cfg, batches[0]['input_ids'], batches[0]['attention_mask'])
    np.testing.assert_allclose(hidden.numpy(), esm_golden['batch0/hidden'], rtol=1e-4, atol=5e-5)
    assert (batches[0]['input_ids'] == cfg.mask_token_id).any(), 'fixture must exercise token dropout'

    got = opool.compute_embeddings(
        batches, lambda b: oesm.esm_forward(sd, cfg, b['input_ids'], b['attention_mask']),
        opool.average_pool)
    np.testing.assert_allclose(got, esm_golden['pooled/mean'], rtol=1e-4, atol=5e-5)


def test_esm_oracle_matches_hf_esm_model(tiny_esm):
    from transformers import EsmModel

    from oracle import esm as oesm

    cfg, sd = tiny_esm
    model = EsmModel(cfg, add_pooling_layer=False)
    model.load_state_dict(sd, strict=False)
    model.eval()
    g = torch.Generator().manual_seed(8)
    ids = torch.randint(4, 24, (3, 70), generator=g)
    ids[:, 0] = 0
    ids[1, 5] = cfg.mask_token_id
    lens = torch.tensor([70, 9, 33])
    mask = (torch.arange(70)[None] < lens[:, None]).long()
    ids = ids.masked_fill(mask == 0, cfg.pad_token_id)
    with torch.no_grad():
        ref = model(input_ids=ids, attention_mask=mask).last_hidden_state
    got = oesm.esm_forward(sd, cfg, ids, mask)
    valid = mask.bool()
    np.testing.assert_allclose(got[valid].numpy(), ref[valid].numpy(), rtol=1e-4, atol=5e-5)


# ---------------------------------------------------------------------------------- Mistral
MISTRAL_CASES = [('full', 'right'), ('full', 'left'), ('window', 'right'), ('window', 'left')]


def _mistral_batches(golden, key):
    for i in range(int(golden['n_batches'])):
        yield {k: torch.from_numpy(golden[f'{key}/batch{i}/{k}']) for k in ('input_ids', 'attention_mask')}


def test_mistral_weights_are_reproducible(mistral_golden):
    from conftest import tiny_mistral_variant

    _, sd = tiny_mistral_variant('full')
    assert weights_digest(sd) == str(mistral_golden['weights_sha256'])


@pytest.mark.parametrize(('variant', 'side'), MISTRAL_CASES)
def test_mistral_oracle_matches_reference(mistral_golden, variant, side):
    """oracle Mistral forward + last-token / mean pool == distllm's AutoEncoder (HF MistralModel)
    outputs: grouped-query causal attention, rotary, RMSNorm, SwiGLU; right and left padding; without
    and with a sliding window; a row truncated at max_position_embeddings."""
    from conftest import tiny_mistral_variant

    from oracle import mistral as omis

    cfg, sd = tiny_mistral_variant(variant)
    key = f'{variant}/{side}'
    batches = list(_mistral_batches(mistral_golden, key))
    assert batches[1]['attention_mask'].sum(1).max() == cfg.max_position_embeddings  # truncated row
    if f'{key}/batch1/hidden' in mistral_golden.files:
        hidden = omis.mistral_forward(sd, cfg, batches[1]['input_ids'], batches[1]['attention_mask'])
        # every position: the reference's SDPA gives zeros on rows that see no key (left padding)
        np.testing.assert_allclose(hidden.numpy(), mistral_golden[f'{key}/batch1/hidden'],
                                   rtol=1e-4, atol=5e-5)

    def encode(b):
        return omis.mistral_forward(sd, cfg, b['input_ids'], b['attention_mask'])

    got = opool.compute_embeddings(batches, encode, opool.last_token_pool)
    np.testing.assert_allclose(got, mistral_golden[f'{key}/pooled/last_token'], rtol=1e-4, atol=5e-5)
    if side == 'right':
        got = opool.compute_embeddings(batches, encode, opool.average_pool, do_normalize=True)
        np.testing.assert_allclose(got, mistral_golden[f'{key}/pooled/mean_normalized'],
                                   rtol=1e-4, atol=5e-5)


def test_mistral_sliding_window_changes_the_result(mistral_golden):
    a = mistral_golden['full/right/pooled/last_token']
    b = mistral_golden['window/right/pooled/last_token']
    long_rows = [1, 6, 8, 9, 10, 11]   # > 80 tokens: the window drops keys
    short_rows = [0, 2, 3, 7]          # shorter than the window: identical
    assert np.abs(a[long_rows] - b[long_rows]).max() > 1e-3
    np.testing.assert_allclose(a[short_rows], b[short_rows], rtol=0, atol=1e-6)


def test_mistral_oracle_matches_hf_model():
    """Independent of the fixtures: the restatement against transformers' MistralModel."""
    from conftest import tiny_mistral_variant
    from transformers import MistralModel

    from oracle import mistral as omis

    cfg, sd = tiny_mistral_variant('window')
    model = MistralModel(cfg)
    model.load_state_dict(sd, strict=False)
    model.eval()
    g = torch.Generator().manual_seed(12)
    ids = torch.randint(4, cfg.vocab_size, (3, 150), generator=g)
    lens = torch.tensor([150, 9, 97])
    mask = (torch.arange(150)[None] < lens[:, None]).long()
    with torch.no_grad():
        ref = model(input_ids=ids, attention_mask=mask).last_hidden_state
    got = omis.mistral_forward(sd, cfg, ids, mask)
    valid = mask.bool()
    np.testing.assert_allclose(got[valid].numpy(), ref[valid].numpy(), rtol=1e-4, atol=5e-5)


def test_modernbert_oracle_matches_reference_and_hf(tiny_modernbert, modernbert_golden):
    """oracle/modernbert.py vs the reference's AutoEncoder outputs (golden) and vs HF ModernBertModel run here:
    rotary per layer type, sliding-window layers (|i - j| <= 16 in the tiny config), GeGLU, layer 0 without
    attn_norm."""
    from transformers import ModernBertModel

    from oracle import modernbert as omb
    from oracle import pooling as opool

    cfg, sd = tiny_modernbert
    g = modernbert_golden
    batches = [{k: torch.from_numpy(g[f'batch{i}/{k}']) for k in ('input_ids', 'attention_mask')}
               for i in range(int(g['n_batches']))]
    hidden = omb.modernbert_forward(sd, cfg, batches[1]['input_ids'], batches[1]['attention_mask'])
    valid = batches[1]['attention_mask'].bool().numpy()
    np.testing.assert_allclose(hidden.numpy()[valid], g['batch1/hidden'][valid], atol=5e-5, rtol=0)
    last = opool.compute_embeddings(batches, lambda b: omb.modernbert_forward(sd, cfg, b['input_ids'], b['attention_mask']),
                                    opool.last_token_pool)
    np.testing.assert_allclose(last, g['pooled/last_token'], atol=5e-5, rtol=0)
    mean = opool.compute_embeddings(
        [{k: v.clone() for k, v in b.items()} for b in batches],
        lambda b: omb.modernbert_forward(sd, cfg, b['input_ids'], b['attention_mask']), opool.average_pool, True)
    np.testing.assert_allclose(mean, g['pooled/mean_normalized'], atol=5e-5, rtol=0)
    # per-layer states against HF's own hidden_states (layer by layer, final_norm applied to the HF states)
    cfg._attn_implementation = 'eager'
    model = ModernBertModel(cfg).eval()
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out = model(input_ids=batches[1]['input_ids'], attention_mask=batches[1]['attention_mask'],
                    output_hidden_states=True)
    states = omb.modernbert_forward(sd, cfg, batches[1]['input_ids'], batches[1]['attention_mask'], return_all=True)
    for layer, ours in enumerate(states, start=1):
        theirs = model.final_norm(out.hidden_states[layer]) if layer < len(states) else out.last_hidden_state
        np.testing.assert_allclose(ours.numpy()[valid], theirs.detach().numpy()[valid], atol=5e-5, rtol=0)
