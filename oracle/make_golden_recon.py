"""Generate tests/golden/recon_tiny.pt from the REAL reference (`utils.chunk_reconstruction.ChunkPTRecon`, imported from
/root/reference) with a recording stand-in for the un-vendored `pytheia` module: every AddView / AddTrack / AddObservation /
SetPoint / SetColor call `create_recon_from_chunk` makes on the committed chunk golden is captured, checked against
oracle.recon_oracle.observation_list and stored. Run in the build container only (needs /root/reference)."""
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Track:
    def __init__(self):
        self.point = None
        self.color = None

    def SetPoint(self, p):
        self.point = np.asarray(p, np.float64).copy()

    def SetColor(self, c):
        self.color = np.asarray(c).copy()

    def SetIsEstimated(self, v):
        pass


class _Reconstruction:
    last = None

    def __init__(self):
        self.views, self.tracks, self.obs = [], [], []
        _Reconstruction.last = self

    def AddView(self, name, group, ts):
        self.views.append(name)
        return len(self.views) - 1

    def MutableView(self, vid):
        return mock.MagicMock()

    def AddTrack(self):
        self.tracks.append(_Track())
        return len(self.tracks) - 1

    def MutableTrack(self, tid):
        return self.tracks[tid]

    def AddObservation(self, view_id, track_id, feature):
        self.obs.append((view_id, track_id, np.asarray(feature, np.float64).copy()))

    def ViewIds(self):
        return list(range(len(self.views)))

    def TrackIds(self):
        return list(range(len(self.tracks)))

    def __getattr__(self, name):          # InitializeInverseDepth etc.
        return mock.MagicMock()


def main():
    pt = mock.MagicMock()
    pt.sfm.Reconstruction = _Reconstruction
    pt.sfm.Feature = lambda xy: np.asarray(xy, np.float64)
    pt.sfm.SetOutlierTracksToUnestimated = lambda *a, **k: 0
    sys.modules["pytheia"] = pt
    sys.modules["pytheia.sfm"] = pt.sfm
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation"):
        sys.modules.setdefault(m, mock.MagicMock())
    # the module is loaded from its file (the reference's utils/__init__ pulls in plyfile, torchcodec, ...); its only
    # other import, pi3.utils.camera.Camera, is a thin pytheia wrapper and is stubbed with the rest of pytheia
    cam_mod = types.ModuleType("pi3.utils.camera")
    cam_mod.Camera = mock.MagicMock
    for name in ("pi3", "pi3.utils"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pi3.utils.camera"] = cam_mod
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_chunk_reconstruction", "/root/reference/utils/chunk_reconstruction.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ChunkPTRecon = mod.ChunkPTRecon
    from oracle.recon_oracle import observation_list

    g = torch.load(os.path.join(ROOT, "tests", "golden", "chunk_tiny.pt"), weights_only=False)
    chunk = dict(g["chunk"])
    # view names play no role in the observation list; the stored 1-tuples trip os.path.basename in the reference (:84)
    chunk["image_paths"] = [[p[0]] for p in chunk["image_paths"]]
    out = {}
    for max_obs in (5, 2):
        rec = ChunkPTRecon()
        rec.set_target_size(chunk["original_width"], chunk["original_height"])
        rec.create_recon_from_chunk(chunk, max_observations_per_track=max_obs)
        r = _Reconstruction.last
        ref = {"track_point": np.stack([t.point for t in r.tracks]), "track_color": np.stack([t.color for t in r.tracks]),
               "obs_view": np.asarray([o[0] for o in r.obs], np.int64), "obs_track": np.asarray([o[1] for o in r.obs], np.int64),
               "obs_xy": np.stack([o[2] for o in r.obs])}
        mine = observation_list(chunk, max_obs)
        for k in ref:
            assert ref[k].shape == mine[k].shape, (k, ref[k].shape, mine[k].shape)
            assert np.array_equal(ref[k], mine[k]), k
        print(f"max_obs {max_obs}: {len(r.tracks)} tracks, {len(r.obs)} observations — restatement identical")
        out[max_obs] = ref
    torch.save({"source": "chunk_tiny.pt['chunk']", "cases": out}, os.path.join(ROOT, "tests", "golden", "recon_tiny.pt"))
    print("wrote tests/golden/recon_tiny.pt")


if __name__ == "__main__":
    main()
