"""Dirichlet index mask and boundary values of the square plate (reference: FEANet/geo.py)."""
import torch


class Geometry():
    """``geometry_idx``: 1.0 on interior nodes, 0.0 on the boundary ring; ``boundary_value``: prescribed values on the
    ring, 0 elsewhere; both (1,1,N,N) fp32 (geo.py:13-30).  The tensors are tagged so that JacobiBlock can recognise
    the default ring without an O(N^2) comparison and use the index-arithmetic fast path (no mask traffic)."""

    def __init__(self, nnode_edge=37, l_shape=False, l_cutout_size=None):
        if l_shape is False:
            self.square_geometry(nnode_edge)
        else:
            self.l_shaped_geometry(nnode_edge, l_cutout_size)

    def square_geometry(self, nnode_edge):
        g = torch.ones(1, 1, nnode_edge, nnode_edge)
        g[0, 0, 0, :] = 0.0
        g[0, 0, -1, :] = 0.0
        g[0, 0, :, 0] = 0.0
        g[0, 0, :, -1] = 0.0
        self.geometry_idx = g
        self.boundary_value = torch.zeros_like(g)
        # fast-path tags for JacobiBlock (index arithmetic instead of mask fields); they carry the tensors' version counters
        # so that an in-place edit made behind set_square_bc's back invalidates them
        self.geometry_idx._mgfea_default_ring = self.geometry_idx._version
        self.boundary_value._mgfea_zero = self.boundary_value._version

    def set_square_bc(self, bc_values):
        '''Input bc is a 2D array, only the locations at boundaries have values'''
        self.boundary_value[:, :, :, :] = bc_values
        self.boundary_value._mgfea_zero = None

    def l_shaped_geometry(self, nnode_edge, l_cutout_size=None):
        # The reference's L-shaped constructor is broken (geo.py:41 unpacks the None returned by square_geometry and
        # raises TypeError); the same error surfaces here instead of inventing semantics.
        raise TypeError("cannot unpack non-iterable NoneType object "
                        "(l_shaped_geometry is broken in the reference: FEANet/geo.py:41)")
