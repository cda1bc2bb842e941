"""Dataset wrappers of the reference (Data/dataset.py:1-104) over its HDF5 files, without h5py / torchvision, and a batch
loader that puts whole batches -- including the per-sample Dirichlet masks of `SingleGrid.ResetBoundary`
(M-FEANet-mg_test.ipynb cell 3) -- on the GPU for the batched V-cycle kernels.

Same class names, constructor arguments, `__len__` / `__getitem__` results as the reference: every item is what
`torchvision.transforms.ToTensor()` makes of an (H, W) array, i.e. a (1, H, W) tensor of the array's dtype.
"""
import numpy as np
import torch
from torch.utils.data import Dataset

from .h5lite import H5File


def _to_tensor(a):
    """ToTensor() on a 2-D numpy array: (H, W) -> (1, H, W), dtype kept (float32 / float64 arrays are not rescaled)"""
    a = np.asarray(a)
    if a.ndim == 3 and a.shape[-1] == 1:  # (H, W, 1) as stored in TestPoisson files: ToTensor moves channels first
        a = a[..., 0]
    return torch.from_numpy(np.ascontiguousarray(a))[None]


class RHSDataSet(Dataset):
    def __init__(self, h5file, case='train', transform=None, target_transform=None):
        """case = 'train' or 'test'"""
        self.data = np.array(H5File(h5file)[case], dtype=np.float32)
        self.transform = transform
        self.target_transform = target_transform

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, idx):
        rhs_tensor = _to_tensor(self.data[idx])
        if self.transform:
            rhs_tensor = self.transform(rhs_tensor)
        return rhs_tensor


class IsoPoissonDataSet(Dataset):
    '''Dataset stores u, f, bc_value, bc_index'''

    def __init__(self, h5file, transform=None, target_transform=None):
        h5 = H5File(h5file)
        self.bc_index = np.array(h5['boundary_index'], dtype=np.float32)
        self.bc_value = np.array(h5['boundary_value'], dtype=np.float32)
        self.f = np.array(h5['rhs'], dtype=np.float32)
        self.u = np.array(h5['u'], dtype=np.float32)
        self.transform = transform
        self.target_transform = target_transform

    def __len__(self):
        return self.f.shape[0]

    def __getitem__(self, idx):
        out = [_to_tensor(a[idx]) for a in (self.u, self.f, self.bc_value, self.bc_index)]
        if self.transform:
            out = [self.transform(t) for t in out]
        return tuple(out)  # u, f, bc_value, bc_index


class IsoPoissonPBCDataSet(Dataset):
    '''Dataset stores f'''

    def __init__(self, h5file, transform=None, target_transform=None):
        self.f = np.array(H5File(h5file)['rhs'], dtype=np.float32)
        self.transform = transform
        self.target_transform = target_transform

    def __len__(self):
        return self.f.shape[0]

    def __getitem__(self, idx):
        f_tensor = _to_tensor(self.f[idx])
        if self.transform:
            f_tensor = self.transform(f_tensor)
        return f_tensor


class TestPoissonDataSet(Dataset):
    __test__ = False  # not a pytest class

    def __init__(self, h5file, transform=None, target_transform=None):
        h5 = H5File(h5file)
        self.dirich_idx = np.array(h5['dirich_idx'], dtype=np.double)
        self.dirich_value = np.array(h5['dirich_value'], dtype=np.double)
        self.traction_idx = np.array(h5['neumann_idx'], dtype=np.double)
        self.traction_value = np.array(h5['neumann_value'], dtype=np.double)
        self.material = np.array(h5['material'], dtype=np.double)
        self.source = np.array(h5['source'], dtype=np.double)
        self.solution = np.array(h5['solution'], dtype=np.double)
        self.transform = transform
        self.target_transform = target_transform

    def __len__(self):
        return self.source.shape[0]

    def __getitem__(self, idx):
        out = [_to_tensor(a[idx]) for a in (self.dirich_idx, self.dirich_value, self.traction_idx, self.traction_value,
                                            self.material, self.source, self.solution)]
        if self.transform:
            out = [self.transform(t) for t in out]
        return tuple(out)


class DeviceBatchLoader:
    """Iterates a map-style dataset in batches and hands every batch over on the GPU: the fields of a batch are stacked in
    pinned host memory (two alternating staging buffers) and copied with one asynchronous H2D copy each, so the batched
    kernels (per-sample Dirichlet masks: `MGTestMultiGrid.forward(u0, F, bc_idx, bc_value, k)`) see (B, 1, N, N) CUDA
    tensors.  The reference feeds `torch.utils.data.DataLoader` batches living on the host (mg_test cell 7)."""

    def __init__(self, dataset, batch_size, shuffle=False, drop_last=False, device=None, seed=None):
        import mgfea

        self.ds, self.bs, self.shuffle, self.drop_last = dataset, int(batch_size), shuffle, drop_last
        self.dev = device or mgfea.require_cuda()
        self.gen = np.random.RandomState(seed)
        self._stage = [None, None]

    def __len__(self):
        n = len(self.ds)
        return n // self.bs if self.drop_last else (n + self.bs - 1) // self.bs

    def __iter__(self):
        order = self.gen.permutation(len(self.ds)) if self.shuffle else np.arange(len(self.ds))
        for bi in range(len(self)):
            idx = order[bi * self.bs:(bi + 1) * self.bs]
            items = [self.ds[int(i)] for i in idx]
            single = torch.is_tensor(items[0])
            fields = [[it] for it in items] if single else [list(it) for it in items]
            nf = len(fields[0])
            slot = bi & 1
            if self._stage[slot] is None or self._stage[slot][0].shape[0] != len(idx):
                self._stage[slot] = [torch.empty((len(idx),) + tuple(fields[0][k].shape), dtype=fields[0][k].dtype,
                                                 pin_memory=True) for k in range(nf)]
            out = []
            for k in range(nf):
                st = self._stage[slot][k]
                for j in range(len(idx)):
                    st[j].copy_(fields[j][k])
                out.append(st.to(self.dev, non_blocking=True))
            torch.cuda.current_stream().synchronize()  # the staging buffer is reused two batches later
            yield out[0] if single else tuple(out)
