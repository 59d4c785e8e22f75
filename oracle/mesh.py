"""oracle/mesh.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of `clean_mesh` (lib/dataset/mesh_util.py:778-791):

    mesh_lst = trimesh.Trimesh(verts, faces).split(only_watertight=False)
    mesh_clean = mesh_lst[argmax(#vertices)]          # first maximum
    return vertices.float(), faces.int()

trimesh is not installable here (PARITY UNPINNED for its internals); what is restated is its published algorithm
(trimesh 3.x, `trimesh/graph.py`): `face_adjacency` = pairs of faces sharing an edge that is used by EXACTLY two
faces (`grouping.group_rows(edges_sorted, require_count=2)`), `split` = connected components of that face graph
(`connected_components(..., min_len=1)` when only_watertight=False), each component re-indexed by `submesh`
(vertices in ascending original id, faces in ascending original index).  Component order follows the smallest face
index of each component (scipy's label order), so "first maximum" = the earliest such component among equals.
"""
import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components


def face_components(faces):
    f = np.asarray(faces, np.int64)
    nf = len(f)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    e.sort(axis=1)
    owner = np.tile(np.arange(nf), 3)
    order = np.lexsort((e[:, 1], e[:, 0]))
    es, os_ = e[order], owner[order]
    same = np.all(es[1:] == es[:-1], axis=1)
    run_start = np.concatenate([[True], ~same])
    run_id = np.cumsum(run_start) - 1
    run_len = np.bincount(run_id)
    pair_first = np.flatnonzero(run_start & (run_len[run_id] == 2))           # edges used by exactly two faces
    a, b = os_[pair_first], os_[pair_first + 1]
    g = coo_matrix((np.ones(len(a)), (a, b)), shape=(nf, nf))
    ncomp, labels = connected_components(g, directed=False)
    return ncomp, labels


def clean_mesh(verts, faces):
    v = np.asarray(verts)
    f = np.asarray(faces, np.int64)
    ncomp, labels = face_components(f)
    best, best_n = -1, -1
    for c in range(ncomp):
        n = len(np.unique(f[labels == c]))
        if n > best_n:
            best, best_n = c, n
    keep_f = f[labels == best]
    ids = np.unique(keep_f)
    remap = np.full(len(v), -1, np.int64)
    remap[ids] = np.arange(len(ids))
    return v[ids].astype(np.float32), remap[keep_f].astype(np.int32)
