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
    with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as f:
        json.dump({"mine": mine, "merged": merged}, f)
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
