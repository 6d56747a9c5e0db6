"""Where does the time of one real chunk go? (creator path, cfg3 chunk shape by default)"""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pi3_slam_b200.pi3 import Pi3
from pi3_slam_b200.synthetic import make_frames, make_state_dict
from pi3_slam_b200.offline_chunk_creator import OfflineChunkCreator, OfflineCreatorConfig
from pi3_slam_b200.host.camera_estimation import estimate_camera_parameters_single_chunk

N, H, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
m = Pi3(); m.load_state_dict(make_state_dict(0)); m = m.cuda().eval()
cfg = OfflineCreatorConfig(model_path="x", output_dir="/tmp/ct_out", device="cuda", do_metric_depth=False,
                           keypoint_type="grid", max_num_keypoints=200, keypoint_seed=1)
cr = OfflineChunkCreator(cfg, model=m)
cr.target_size = (H, W)
imgs = make_frames(N, H, W)[None].pin_memory()
paths = [(f"f{i}.png",) for i in range(N)]
for _ in range(2):
    out = cr._process_single_chunk(imgs, paths)
torch.cuda.synchronize()
t0 = time.perf_counter(); out = cr._process_single_chunk(imgs, paths); torch.cuda.synchronize(); t_all = time.perf_counter() - t0
res = m(imgs.cuda()); torch.cuda.synchronize()
t0 = time.perf_counter(); res = m(imgs.cuda()); torch.cuda.synchronize(); t_fwd = time.perf_counter() - t0
t0 = time.perf_counter(); cp = estimate_camera_parameters_single_chunk(res["local_points"][0], res["conf"][0]); torch.cuda.synchronize(); t_intr = time.perf_counter() - t0
t0 = time.perf_counter(); torch.save(out, "/tmp/ct_out/x.pt"); t_save = time.perf_counter() - t0
print(json.dumps({"frames": N, "size": [H, W], "chunk_total_s": t_all, "forward_s": t_fwd, "intrinsics_scipy_s": t_intr,
                  "torch_save_s": t_save, "other_s": t_all - t_fwd - t_intr, "frames_per_s_chunk": N / t_all,
                  "cpu_threads": torch.get_num_threads()}))
