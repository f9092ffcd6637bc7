"""Host-side view of the device neighbour search (reference ``torchmd/neighbourlist.py``).

The reference module is dead code: ``discretize_box`` (``neighbourlist.py:4-31``) is never
called, its fixed-width last bin would miss pairs across the periodic boundary when the box
is not a multiple of the cell size, and ``neighbour_list`` is commented out
(``neighbourlist.py:34-47``).  What the reference actually uses as its neighbour list is
``ava_idx[dist <= cutoff]`` (``forces.py:76-81,264-269``).  This module exposes the
replacement that runs on the device (``csrc/neighbor.cuh``):

* ``cell_grid`` -- the cell grid the library lays over a box (same arithmetic as
  ``finalize()`` in ``csrc/tmd_b200.cu``): cells of equal width ``L/n >= (cutoff+skin)/2``,
  a dimension too short for five cells collapses to one cell that spans it;
* ``neighbour_list`` -- the pairs inside the cutoff, exactly the reference's set.
"""
import math

LIST_MARGIN = 0.004  # A, slack of the approximate list-build arithmetic (csrc/tmd_b200.cu)
CELLS_PER_RADIUS = 2  # nsub


def cell_grid(box_lengths, cutoff, skin=1.0):
    """Cells per dimension and neighbour-cell reach for a periodic orthorhombic box.

    Unlike ``discretize_box`` the cells tile the box exactly (width ``L/n``), and the sweep
    covers ``2*reach+1`` cells per dimension with periodic wrap, so no pair within
    ``cutoff+skin`` can be missed.
    """
    rlist = cutoff + skin + LIST_MARGIN
    ncell, reach, width = [], [], []
    for L in box_lengths:
        n = min(int(math.floor(L * CELLS_PER_RADIUS / rlist)), 128)
        if n >= 2 * CELLS_PER_RADIUS + 1:
            ncell.append(n)
            reach.append(CELLS_PER_RADIUS)
            width.append(L / n)
        else:  # one cell spans the dimension; the minimum image is applied per pair instead
            ncell.append(1)
            reach.append(0)
            width.append(L)
    return {"ncell": tuple(ncell), "reach": tuple(reach), "width": tuple(width), "rlist": rlist}


def neighbour_list(forces, pos, box, replica=0):
    """(P,2) int32 CUDA tensor of the pairs ``i<j`` with ``dist <= cutoff`` for one replica,
    lexicographically sorted: bit for bit the reference's ``ava_idx[dist <= cutoff]``.
    ``forces`` is a ``torchmd_b200.Forces`` (it owns the exclusions and the cutoff)."""
    return forces.neighbour_pairs(pos, box, replica=replica)
