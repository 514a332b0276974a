"""Overlap / micro-batch samplers against batch sequences recorded from the
reference's own sampler classes (tests/golden/samplers.json, tests/golden/gen_golden.py)."""
import json
import os

import pytest

from dd4ml_b200.dataloaders import MicroBatchFlattenSampler, MicroBatchOverlapSampler, OverlapBatchSampler

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "samplers.json")))["cases"]


@pytest.mark.parametrize("c", CASES, ids=lambda c: f"n{c['n']}_b{c['batch_size']}_o{c['overlap']}_d{int(c['drop_last'])}_s{c['num_subdomains']}")
def test_batch_sequences_match_reference(c):
    ob = OverlapBatchSampler(c["order"], c["batch_size"], overlap=c["overlap"], drop_last=c["drop_last"])
    assert ob.overlap_sz == c["overlap_sz"]
    assert len(ob) == c["len"]
    assert [list(b) for b in ob] == c["batches"]
    assert [list(b) for b in ob] == c["batches"]          # second epoch: the base order is cached
    if c["num_subdomains"] > 0 and c["overlap_sz"] > 0:
        if "micro_error" in c:
            exc = {"RuntimeError": RuntimeError, "ValueError": ValueError}[c["micro_error"]]
            with pytest.raises(exc):
                mb = MicroBatchOverlapSampler(ob, c["num_subdomains"], allow_empty_microbatches=False)
                list(mb)
        else:
            mb = MicroBatchOverlapSampler(ob, c["num_subdomains"], allow_empty_microbatches=False)
            assert [[list(p) for p in parts] for parts in mb] == c["micro"]
            fl = MicroBatchFlattenSampler(mb)
            assert [list(p) for p in fl] == c["flat"]
            assert len(fl) == c["flat_len"]
            assert mb.get_overlap_info() == c["info"]


def test_constructor_errors_match_reference():
    ob0 = OverlapBatchSampler(range(10), 4, overlap=0.0)
    with pytest.raises(ValueError):
        MicroBatchOverlapSampler(ob0, 2)                   # overlap == 0
    ob = OverlapBatchSampler(range(10), 4, overlap=2)
    with pytest.raises(ValueError):
        MicroBatchOverlapSampler(ob, 0)
    with pytest.raises(ValueError):
        MicroBatchOverlapSampler(ob, 7)                    # more subdomains than batch + overlap
    with pytest.raises(TypeError):
        MicroBatchOverlapSampler([[0, 1]], 2)
    assert OverlapBatchSampler(range(10), 4, overlap=9).overlap_sz == 3      # int overlap is capped at bs - 1
    assert OverlapBatchSampler(range(10), 4, overlap=1.5).overlap_sz == 0    # a float >= 1 means no overlap
    assert list(OverlapBatchSampler([], 4, overlap=1)) == []
    mb = MicroBatchOverlapSampler(ob, 7, allow_empty_microbatches=True)
    assert any(len(p) == 0 for parts in mb for p in parts)


# ------------------------------------------------------------------ property test against a naive restatement
def _naive_batches(order, bs, ov_sz, drop_last):
    """Element-by-element restatement of dataloaders.py:57-96 (loops, modular indexing)."""
    n = len(order)
    if n == 0:
        return []
    nb = n // bs if drop_last else -(-n // bs)
    out = []
    for b in range(nb):
        start, end = b * bs, min(b * bs + bs, n)
        uniq = [order[i] for i in range(start, end)]
        if not drop_last and end >= n and len(uniq) < bs:
            out.append(uniq)
            continue
        if ov_sz > 0:
            if drop_last and b == nb - 1:
                extra = [order[i] for i in range(ov_sz)]
            else:
                extra = [order[(end % n + i) % n] for i in range(ov_sz)]
            uniq = uniq + extra
        out.append(uniq)
    return out


def test_overlap_sampler_matches_naive_restatement_on_random_cases():
    import random
    rnd = random.Random(11)
    for _ in range(300):
        n = rnd.randint(0, 60)
        bs = rnd.randint(1, 20)
        overlap = rnd.choice([0, 0.0, rnd.random(), rnd.randint(1, 25), 1.5])
        drop = rnd.random() < 0.5
        order = list(range(n))
        rnd.shuffle(order)
        frac = isinstance(overlap, float) and 0.0 < overlap < 1.0
        if frac and int(round(overlap * bs)) >= bs:       # dataloaders.py:44-45 refuses it
            with pytest.raises(ValueError):
                OverlapBatchSampler(order, bs, overlap=overlap, drop_last=drop)
            continue
        ob = OverlapBatchSampler(order, bs, overlap=overlap, drop_last=drop)
        assert 0 <= ob.overlap_sz < bs
        got = [list(b) for b in ob]
        assert got == _naive_batches(order, bs, ob.overlap_sz, drop), (n, bs, overlap, drop)
        assert len(got) == len(ob)
        if ob.overlap_sz > 0 and n >= bs:
            nsub = rnd.randint(1, bs + ob.overlap_sz)
            mb = MicroBatchOverlapSampler(ob, nsub, allow_empty_microbatches=True)
            for parts, batch in zip(mb, got):
                nu = min(len(batch), bs)
                assert sorted(x for p in parts for x in p) == sorted(batch)          # a partition of the mini-batch
                assert all(p == batch[:nu][j::nsub] + batch[nu:][j::nsub] for j, p in enumerate(parts))
