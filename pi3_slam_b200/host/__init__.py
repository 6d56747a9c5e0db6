"""Host-side (Python) mirrors of the reference's small pipeline utilities on the chunk path:
index bookkeeping, target size, grid keypoints, intrinsics estimation. No heavy math lives here."""
