# Drop-in names of the reference's lib/dataset/mesh_util.py that belong to the accelerated path.
from icon_b200.visibility import get_visibility  # noqa: F401  (reference: lib/dataset/mesh_util.py:280-316)
from icon_b200.voxelize import read_smpl_constants  # noqa: F401  (reference: lib/dataset/mesh_util.py:240-263)
