"""One rank of the world_size-2 `gloo` tests (tests/test_sharding_cpu.py starts two of these).

The host logic under test is gecoz_b200/sharding.py; the compute behind it is injected: here the CPU oracle
(test infrastructure), on the GPU box the CUDA engine.  Usage:
    python tests/dist_worker.py RANK WORLD INIT_FILE OUT_DIR build|query
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def records():
    from gecoz_b200 import synth
    lens = [40_000, 26_000, 15_000, 14_000, 9_000, 2_500, 600, 600, 31]
    return [(f"seq{i} test", synth.iid_acgtn(ln, 70 + i)) for i, ln in enumerate(lens)]


class OracleEngine:
    """BlockWriter.run through the oracle: the stand-in for the CUDA engine when no GPU exists."""

    def symbol_counts(self, text):
        return np.bincount(text, minlength=256).astype(np.int64)

    def build_block(self, text, n, sampling_rate, shape, gcz_out, gcx_out):
        from oracle import gcz_oracle as O
        r = O.build_block(text, sampling_rate)
        gcz_out[:] = r["gcz_body"]
        gcx_out[:] = r["gcx_body"]
        return {}


class OracleGSSA:
    """The slice of the GSSA interface sharding.py uses, answered by the oracle."""

    def __init__(self, text):
        from oracle import gcz_oracle as O
        r = O.build_block(text, 32)
        self.g = O.GSSA(r["gcz_body"], len(text), r["gcx_body"])
        self.n_strings = self.g.n_strings

    def count_batch(self, packed):
        sp, ep, _ = self.g.search_batch(*packed)
        return sp, ep

    def find_batch_raw(self, packed):
        data, off = packed
        n = len(off) - 1
        per = np.zeros((n, self.n_strings), np.int64)
        pos, poff = [], np.zeros(n + 1, np.int64)
        for i in range(n):
            res = self.g.find(bytes(data[off[i]:off[i + 1]]))
            if res is not None:
                for s, a in enumerate(res):
                    if a is not None:
                        per[i, s] = len(a)
                        pos.append(a)
            poff[i + 1] = poff[i] + per[i].sum()
        return per, (np.concatenate(pos) if pos else np.zeros(0, np.int64)), poff


def query_inputs():
    from gecoz_b200 import synth
    recs = records()
    texts = [synth.block_of([recs[0][1]]), synth.block_of([recs[1][1], recs[4][1], recs[8][1]])]
    data, off = synth.patterns(texts[0], 301, 4, 12, seed=9)          # 301: shards of unequal size
    return texts, data, off


def main():
    rank, world, init_file, out_dir, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], Path(sys.argv[4]), sys.argv[5]
    import torch.distributed as dist
    from gecoz_b200 import sharding
    gpu = mode.endswith("_gpu")
    device = None
    if gpu:                                              # the real thing: NCCL, one GPU per rank, CUDA engine
        import torch
        import gecoz_b200 as G
        torch.cuda.set_device(rank)
        device = f"cuda:{rank}"
        dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world,
                                device_id=torch.device(device))
    else:
        dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        if mode == "build_gpu":
            info = sharding.sharded_index_records(records(), out_dir / "x.gcz", rank=rank, world=world,
                                                  engine=sharding.GpuEngine(rank))
            (out_dir / f"mine{rank}.txt").write_text(" ".join(map(str, info["mine"])))
        elif mode == "query_gpu":
            from oracle import gcz_oracle as O
            texts, data, off = query_inputs()
            gssas = []
            for t in texts:
                r = O.build_block(t, 32)                 # the index files are given; the queries are what is tested
                gssas.append(G.GSSA.open(rank, r["gcz_body"], len(t), r["gcx_body"]))
            c = sharding.count_sharded(gssas, data, off, rank=rank, world=world, device=device)
            f = sharding.find_sharded(gssas, data, off, rank=rank, world=world, device=device)
            tot = sharding.count_totals_sharded(gssas, data, off, rank=rank, world=world, device=device)
            if rank == 0:
                assert np.array_equal(tot, np.maximum(c[1] - c[0] + 1, 0).sum(axis=0))
                np.savez(out_dir / "query.npz", sp=c[0], ep=c[1],
                         **{f"per{b}": x[0] for b, x in enumerate(f)}, **{f"pos{b}": x[1] for b, x in enumerate(f)},
                         **{f"off{b}": x[2] for b, x in enumerate(f)})
        elif mode == "build":
            info = sharding.sharded_index_records(records(), out_dir / "x.gcz", rank=rank, world=world, engine=OracleEngine())
            (out_dir / f"mine{rank}.txt").write_text(" ".join(map(str, info["mine"])))
        else:
            texts, data, off = query_inputs()
            gssas = [OracleGSSA(t) for t in texts]
            c = sharding.count_sharded(gssas, data, off, rank=rank, world=world)
            f = sharding.find_sharded(gssas, data, off, rank=rank, world=world)
            tot = sharding.count_totals_sharded(gssas, data, off, rank=rank, world=world)
            # the gather the CUDA paths use (tensors stay where they are: here on the CPU), shards of unequal and of zero length
            import torch
            for lens in ([3, 5], [0, 4], [2, 0], [0, 0]):
                parts = sharding._gather_device(torch.arange(lens[rank], dtype=torch.int64) + 100 * rank, rank=rank, world=world)
                if rank == 0:
                    assert [p.tolist() for p in parts] == [list(range(100 * r, 100 * r + lens[r])) for r in range(world)]
                else:
                    assert parts is None
            if rank == 0:
                assert np.array_equal(tot, np.maximum(c[1] - c[0] + 1, 0).sum(axis=0))
                np.savez(out_dir / "query.npz", sp=c[0], ep=c[1],
                         **{f"per{b}": x[0] for b, x in enumerate(f)}, **{f"pos{b}": x[1] for b, x in enumerate(f)},
                         **{f"off{b}": x[2] for b, x in enumerate(f)})
            else:
                assert c is None and f is None
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
