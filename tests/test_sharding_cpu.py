"""Multi-GPU host logic (gecoz_b200/sharding.py) on CPU: pure partitioning functions, and world_size-2 `gloo`
runs of the sharded build and the sharded queries with the oracle injected as the compute engine."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

HERE = Path(__file__).resolve().parent


def test_lpt_assign_hg38_blocks():
    from gecoz_b200 import sharding, synth
    from gecoz_b200.geco_index import FastaSequence, merge_blocks
    blocks = merge_blocks([FastaSequence(h, n) for h, n in zip(synth.HG38_NAMES, synth.HG38_LENGTHS)])
    sizes = [b.size for b in blocks]
    assert len(sizes) == 18 and sum(sizes) == 3_088_286_426
    for world, min_eff in ((1, 1.0), (2, 0.99), (4, 0.90), (8, 0.80)):
        owner = sharding.lpt_assign(sizes, world)
        load = [sum(s for s, r in zip(sizes, owner) if r == k) for k in range(world)]
        assert sorted(set(owner)) == list(range(world))
        assert sum(load) == sum(sizes)
        assert sum(sizes) / world / max(load) >= min_eff                     # LPT efficiency (App. D: 1.00/0.996/0.92/0.81)
    assert sharding.lpt_assign([5, 5, 5], 2) == [0, 1, 0]                    # ties: file order, lowest rank
    assert sharding.lpt_assign([], 3) == []


def test_shard_bounds():
    from gecoz_b200 import sharding
    assert sharding.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sharding.shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert sharding.shard_bounds(0, 2) == [(0, 0), (0, 0)]
    for n in (1, 7, 100, 301):
        for w in (1, 2, 3, 8):
            b = sharding.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(x[1] == y[0] for x, y in zip(b, b[1:]))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_single_rank_needs_no_process_group(tmp_path):
    """world == 1 never touches torch.distributed."""
    sys.path.insert(0, str(HERE))
    import dist_worker as W
    from gecoz_b200 import sharding
    from oracle import gcz_oracle as O
    recs = W.records()
    sharding.sharded_index_records(recs, tmp_path / "x.gcz", engine=W.OracleEngine())
    gcz, gcx, _ = O.write_files([(h, s.tobytes()) for h, s in recs])
    assert (tmp_path / "x.gcz").read_bytes() == gcz and (tmp_path / "x.gcx").read_bytes() == gcx


def _run_world(tmp_path, mode, world=2):
    init = tmp_path / f"rendezvous_{mode}"
    procs = [subprocess.Popen([sys.executable, str(HERE / "dist_worker.py"), str(r), str(world), str(init), str(tmp_path), mode],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    for p in procs:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, out


def test_sharded_build_gloo_world2(tmp_path):
    """Two ranks write one .gcz/.gcx pair: byte-identical to the single-process writer, every block built once."""
    sys.path.insert(0, str(HERE))
    import dist_worker as W
    from oracle import gcz_oracle as O
    _run_world(tmp_path, "build")
    recs = W.records()
    gcz, gcx, blocks = O.write_files([(h, s.tobytes()) for h, s in recs])
    assert (tmp_path / "x.gcz").read_bytes() == gcz
    assert (tmp_path / "x.gcx").read_bytes() == gcx
    mine = [list(map(int, (tmp_path / f"mine{r}.txt").read_text().split())) for r in range(2)]
    assert sorted(mine[0] + mine[1]) == list(range(len(blocks))) and mine[0] and mine[1]


def test_sharded_queries_gloo_world2(tmp_path):
    """Pattern batch cut in two, results gathered on rank 0: equal to the unsharded answers."""
    sys.path.insert(0, str(HERE))
    import dist_worker as W
    _run_world(tmp_path, "query")
    got = np.load(tmp_path / "query.npz")
    texts, data, off = W.query_inputs()
    for b, t in enumerate(texts):
        g = W.OracleGSSA(t)
        sp, ep = g.count_batch((data, off))
        assert np.array_equal(got["sp"][b], sp) and np.array_equal(got["ep"][b], ep)
        per, pos, poff = g.find_batch_raw((data, off))
        assert np.array_equal(got[f"per{b}"], per)
        assert np.array_equal(got[f"pos{b}"], pos)
        assert np.array_equal(got[f"off{b}"], poff)
    assert int((got["ep"] >= got["sp"]).sum()) > 100


class _OracleFileEngine:
    """Engine for gecoz_file.GecozFileWriter (device passed per call) backed by the oracle."""

    def __init__(self, fail_first=False):
        self.fail_first, self.builds = fail_first, 0

    def symbol_counts(self, text, device):
        return np.bincount(np.asarray(text), minlength=256).astype(np.int64)

    def build_block(self, device, text, n, sampling_rate, shape, gcz_out, gcx_out):
        from gecoz_b200 import _native as N
        from oracle import gcz_oracle as O
        self.builds += 1
        if self.fail_first and self.builds == 1:
            raise N.GczOutOfMemory(N.GCZ_E_NOMEM, "injected")
        r = O.build_block(np.asarray(text), sampling_rate)
        assert len(r["gcz_body"]) == len(gcz_out) and len(r["gcx_body"]) == len(gcx_out)
        gcz_out[:] = r["gcz_body"]
        gcx_out[:] = r["gcx_body"]
        return {"total_ms": 0.0}


@pytest.mark.parametrize("devices,fail_first", [((0,), False), ((0, 1, 2), False), ((0,), True)])
def test_python_file_writer_with_the_oracle_engine(tmp_path, devices, fail_first):
    """gecoz_file.GecozFileWriter (pooled block buffers, header + body written with one pwrite per file after the device
    token is released) produces the oracle's files, with several blocks in flight and across the out-of-memory retry."""
    from gecoz_b200 import synth
    from gecoz_b200.geco_index import index_records
    from gecoz_b200.gecoz_file import GecozFileWriter
    from oracle import gcz_oracle as O
    recs = [(f"r{i} d", synth.iid_acgtn(int(ln), 60 + i)) for i, ln in enumerate([7000, 6900, 6800, 3000, 2900, 800, 40, 0])]
    eng = _OracleFileEngine(fail_first)
    import gecoz_b200.geco_index as GI
    real = GI.GecozFileWriter
    GI.GecozFileWriter = lambda o, x, s, d: GecozFileWriter(o, x, s, d, engine=eng)
    try:
        info = index_records(recs, tmp_path / "p.gcz", sampling=16, devices=devices)
    finally:
        GI.GecozFileWriter = real
    gcz, gcx, blocks = O.write_files([(h, s.tobytes()) for h, s in recs], 16)
    assert (tmp_path / "p.gcz").read_bytes() == gcz and (tmp_path / "p.gcx").read_bytes() == gcx
    assert len(info["blocks"]) == len(blocks) and eng.builds == len(blocks) + (1 if fail_first else 0)
