"""CPU, world_size 2 over gloo: the host-side logic of the chunk-sharded path (ownership + manifest gather)."""
import json
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pi3_slam_b200.host.datasets import chunk_index_list, shard_chunks
    from pi3_slam_b200.offline_chunk_creator import OfflineChunkCreator
    chunks = chunk_index_list(1000, 100, 20)
    mine = shard_chunks(len(chunks), rank, world)
    manifest = [{"chunk_index": i, "file": f"chunk_{i:06d}.pt", "start_idx": chunks[i][0], "end_idx": chunks[i][1],
                 "num_frames": chunks[i][1] - chunks[i][0], "image_paths": [[f"f{j}"] for j in range(*chunks[i])]}
                for i in mine]
    merged = OfflineChunkCreator._gather_manifest(manifest)
    # hybrid schedule: the groups of the sequence-parallel tail wave are created collectively, in plan order
    from pi3_slam_b200.host.datasets import plan_waves
    waves = plan_waves([e - b for b, e in chunks], world)
    tail = waves[-1]
    assert tail == [(12, (0, 1))]
    group = dist.new_group(ranks=list(tail[0][1]))
    part = torch.full((2, 3), float(rank))
    full = torch.empty(4, 3)
    dist.all_gather_into_tensor(full, part, group=group)       # frame-ordered gather of per-frame outputs
    assert full[:2].eq(0).all() and full[2:].eq(1).all()
    # a chunk shared by a group is entered by all of its ranks or by none
    from pi3_slam_b200.host.datasets import group_all_ok
    agree = [group_all_ok(True, group), group_all_ok(rank != 1, group), group_all_ok(False, group)]
    with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as f:
        json.dump({"mine": mine, "merged": merged, "agree": agree}, f)
    dist.destroy_process_group()


def test_chunk_sharding_and_manifest_gather_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = json.load(open(tmp_path / "rank0.json"))
    r1 = json.load(open(tmp_path / "rank1.json"))
    assert r0["mine"] == [0, 2, 4, 6, 8, 10, 12] and r1["mine"] == [1, 3, 5, 7, 9, 11]
    assert r0["merged"] == r1["merged"]
    assert [e["chunk_index"] for e in r0["merged"]] == list(range(13))
    assert r0["merged"][12]["start_idx"] == 960 and r0["merged"][12]["num_frames"] == 40
    assert r0["agree"] == [True, False, False] and r1["agree"] == [True, False, False]


def test_plan_waves_properties():
    from pi3_slam_b200.host.datasets import chunk_index_list, plan_waves, shard_chunks
    frames = [e - b for b, e in chunk_index_list(1000, 100, 20)]           # 12 x 100 + 40
    for world in (1, 2, 4, 8):
        waves = plan_waves(frames, world)
        seen = [c for w in waves for c, _ in w]
        assert seen == list(range(13))                                     # every chunk exactly once, in order
        for w in waves:
            used = [r for _, ranks in w for r in ranks]
            assert len(used) == len(set(used)) and all(0 <= r < world for r in used)
            assert all(len(ranks) <= frames[c] for c, ranks in w)
        # plain sharding when sequence parallelism is off
        flat = plan_waves(frames, world, allow_sp=False)
        for r in range(world):
            assert [c for w in flat for c, ranks in w if ranks == (r,)] == shard_chunks(13, r, world)
    w8 = plan_waves(frames, 8)
    assert [len(w) for w in w8] == [8, 4, 1] and w8[1][0] == (8, (0, 1)) and w8[2][0][1] == tuple(range(8))
    # makespan in chunk-times (cost ~ frames^2 for the dominant global attention; SP divides it by the group size)
    cost = lambda wv: max((frames[c] / 100) ** 2 / len(r) for c, r in wv)          # noqa: E731
    assert sum(cost(w) for w in w8) < 1.53 and sum(cost(w) for w in plan_waves(frames, 8, False)) == 2.0
    # a frame count that does not divide is split unevenly (19 + 18 frames), the group keeps its size
    from pi3_slam_b200.host.datasets import split_frames
    assert plan_waves([100] * 5 + [37], 4)[-1] == [(4, (0, 1)), (5, (2, 3))]
    assert split_frames(37, 2) == [(0, 19), (19, 18)] and split_frames(300, 8)[3:5] == [(114, 38), (152, 37)]
    assert sum(n for _, n in split_frames(300, 8)) == 300 and split_frames(3, 2) == [(0, 2), (2, 1)]
    # fewer frames than ranks: the group shrinks
    assert plan_waves([2], 4) == [[(0, (0, 1))]]


def test_sequence_parallel_gather_plan_compacts_remote_rows(monkeypatch):
    """Pi3._sp_plan: the row moves that make every OTHER rank's K|V rows contiguous in the gathered (padded) buffer —
    even and uneven frame splits, every rank position. Pure host logic: emulated on a CPU tensor."""
    import torch
    import torch.distributed as dist

    from pi3_slam_b200.pi3 import Pi3

    for counts in ([2, 2], [38, 38, 38, 38, 37, 37, 37, 37], [3, 2, 2], [1, 1, 1, 1], [5, 4], [2, 2, 2, 1], [4, 4, 3, 3, 3, 3, 3, 3]):
        for r in range(len(counts)):
            monkeypatch.setattr(dist, "get_world_size", lambda group=None, n=len(counts): n)
            monkeypatch.setattr(dist, "get_rank", lambda group=None, rr=r: rr)
            T = 3
            sp = Pi3._sp_plan(None, counts, T)
            m = sp["m_max"]
            assert m == max(counts) * T and sp["all_rows"] == sum(counts) * T
            buf = torch.full((len(counts) * m,), -1, dtype=torch.int64)
            for g, c in enumerate(counts):
                buf[g * m: g * m + c * T] = g * 1000 + torch.arange(c * T)
            for dst, src, n in sp["copies"]:
                assert dst + n <= src or src + n <= dst                # moves never overlap
                buf[dst:dst + n] = buf[src:src + n].clone()
            want = sorted(g * 1000 + i for g, c in enumerate(counts) if g != r for i in range(c * T))
            assert sorted(buf[: sp["remote_rows"]].tolist()) == want, (counts, r)
