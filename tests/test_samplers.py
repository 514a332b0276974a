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
