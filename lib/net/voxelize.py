from icon_b200.voxelize import Voxelization, read_smpl_constants  # noqa: F401  (reference: lib/net/voxelize.py:66-137)
