"""B200-native drop-in for the hot path of longfish/Multigrid-FEANet (``FEANet`` package of the reference).

Same module / class / method names as the reference (FEANet/{mesh,geo,model,jacobi,multigrid}.py) plus the notebook
drivers promoted to importable code (``FEANet.drivers``).  Tensors live on the GPU; every operator is a hand-written
sm_100a kernel reached through the C ABI of ``libmgfea.so`` (``mgfea`` package).  No CPU fallback.
"""
from . import geo, jacobi, mesh, model  # noqa: F401

__all__ = ["geo", "jacobi", "mesh", "model", "multigrid", "drivers", "solver"]
