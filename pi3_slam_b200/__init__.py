"""pi3_slam_b200 — B200-native (sm_100a) implementation of Pi3_SLAM's offline chunk-creation hot path.

Public surface mirrors the reference for this path only:
  * ``Pi3``                         (reference: pi3/models/pi3.py)
  * ``OfflineChunkCreator`` / ``OfflineCreatorConfig``   (reference: slam/offline_chunk_creator.py)
All heavy math runs in ``libpi3b200.so`` (hand-written CUDA behind the C ABI in include/pi3b200.h).
"""
__version__ = "0.1.0"
