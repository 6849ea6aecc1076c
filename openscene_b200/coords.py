"""Host side of the coordinate manager: owns the per-tensor-stride coordinate sets, their hash
tables and the kernel-map cache, all built by ``libosb200`` (csrc/coords.cu).

Mirrors what MinkowskiEngine's CoordinateManager does for the reference:
one forward of MinkUNet builds the stride-1 set, four coarser sets and 10 kernel maps, each once.
"""
import ctypes

import torch

from . import _cabi as C


def _next_pow2(n):
    p = 1
    while p < n:
        p <<= 1
    return p


class CoordSet:
    """One coordinate set (internal Morton order) + its lazily built hash table."""
    __slots__ = ('coords', 'n', 'slots', 'cap', 'ts')

    def __init__(self, coords, ts):
        self.coords, self.n, self.ts = coords, coords.shape[0], ts
        self.slots, self.cap = None, 0

    def ensure_hash(self):
        if self.slots is None:
            self.cap = _next_pow2(max(4 * self.n, 16))      # load factor <= 0.25: short probe chains (the longest of 32 lanes gates a warp)
            self.slots = torch.empty(self.cap * 16, dtype=torch.uint8, device=self.coords.device)
            C.call('osb_hash_build', C.ptr(self.coords), self.n, C.ptr(self.slots), self.cap, C.stream_ptr())
        return self


class KernelMap:
    """Output-stationary map: nbr[k, o] = input row feeding output row o through offset k (-1: none)."""
    __slots__ = ('nbr', 'K', 'n_in', 'n_out', 'pairs_per_k', '_t')

    def __init__(self, nbr, K, n_in, n_out, pairs_per_k=None):
        self.nbr, self.K, self.n_in, self.n_out, self.pairs_per_k = nbr, K, n_in, n_out, pairs_per_k
        self._t = None

    def transposed(self):
        """Map with input/output roles swapped (transposed conv forward, conv dgrad)."""
        if self._t is None:
            nbr_t = torch.empty((self.K, self.n_in), dtype=torch.int32, device=self.nbr.device)
            C.call('osb_kernel_map_transpose', C.ptr(self.nbr), self.n_out, self.K, C.ptr(nbr_t), self.n_in,
                   C.stream_ptr())
            # no back-reference from the transposed map: a reference cycle would keep both device buffers alive until the
            # cyclic garbage collector runs (never, inside a gc-disabled serving / benchmark loop)
            self._t = KernelMap(nbr_t, self.K, self.n_out, self.n_in)
        return self._t

    def num_pairs(self):
        if self.pairs_per_k is None:
            return int((self.nbr >= 0).sum().item())
        return int(self.pairs_per_k.sum().item())


class CoordinateManager:
    def __init__(self, coordinates, pyramid_levels=0):
        """coordinates: int32 CUDA tensor [N,4] = (batch, x, y, z), unique rows, caller order.
        pyramid_levels > 0 builds the tensor-stride 2, 4, ... sets in the same native call (one host sync for the
        whole encoder pyramid instead of two per level)."""
        C.require_cuda(coordinates, 'coordinates')
        coords = coordinates.to(torch.int32).contiguous()
        assert coords.dim() == 2 and coords.shape[1] == 4, "coordinates must be [N,4] (batch,x,y,z)"
        n = coords.shape[0]
        if n == 0:
            raise RuntimeError("openscene_b200: empty coordinate set")
        dev = coords.device
        self.device = dev
        with torch.cuda.device(dev):
            coords_int = torch.empty_like(coords)
            self.perm = torch.empty(n, dtype=torch.int32, device=dev)
            self.inv_perm = torch.empty(n, dtype=torch.int32, device=dev)
            cs = CoordSet(coords_int, 1)
            cs.cap = _next_pow2(max(4 * n, 16))
            cs.slots = torch.empty(cs.cap * 16, dtype=torch.uint8, device=dev)
            ws_bytes = C.lib().osb_coordset_workspace_bytes(n)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            status = (ctypes.c_int32 * 2)(0, 0)
            pyr = None
            if pyramid_levels > 0:
                L = pyramid_levels
                lvl = torch.empty((L, n, 4), dtype=torch.int32, device=dev)
                par = torch.empty((L, n), dtype=torch.int32, device=dev)
                counts = (ctypes.c_int64 * (L + 1))()
                C.call('osb_coordset_pyramid', C.ptr(coords), n, L, C.ptr(coords_int), C.ptr(self.perm), C.ptr(self.inv_perm),
                       C.ptr(cs.slots), cs.cap, C.ptr(lvl), C.ptr(par), counts, status, C.ptr(ws), ws_bytes, C.stream_ptr())
                pyr = (lvl, par, list(counts))
            else:
                C.call('osb_coordset_build', C.ptr(coords), n, C.ptr(coords_int), C.ptr(self.perm), C.ptr(self.inv_perm),
                       C.ptr(cs.slots), cs.cap, status, C.ptr(ws), ws_bytes, C.stream_ptr())
        if status[0] & 1:
            raise RuntimeError("openscene_b200: coordinate out of range (need 0 <= batch < 1024, |x|,|y|,|z| < 2^17-256)")
        if status[0] & 2:
            raise RuntimeError("openscene_b200: duplicate coordinates in SparseTensor input; the reference's loaders "
                               "always voxelise first (dataset/voxelizer.py:128), so rows are unique per scene")
        self.sets = {1: cs}
        self.parent_of = {}
        self.kmaps = {}
        self._ws = ws
        if pyr is not None:
            lvl, par, counts = pyr
            ts = 1
            for l in range(len(counts) - 1):
                self.sets[2 * ts] = CoordSet(lvl[l, :counts[l + 1]], 2 * ts)
                self.parent_of[(ts, 2 * ts)] = par[l, :counts[l]]
                ts *= 2

    # -- coordinate sets -------------------------------------------------------------------
    def stride(self, ts, s):
        """Tensor stride ts -> ts*s: unique(floor(c/(ts*s))*(ts*s)); cached."""
        new = ts * s
        if new not in self.sets:
            fine = self.sets[ts]
            with torch.cuda.device(self.device):
                coarse = torch.empty_like(fine.coords)
                parent = torch.empty(fine.n, dtype=torch.int32, device=self.device)
                n_c = ctypes.c_int64(0)
                ws_bytes = C.lib().osb_coordset_workspace_bytes(fine.n)
                if self._ws.numel() < ws_bytes:
                    self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
                C.call('osb_coordset_stride', C.ptr(fine.coords), fine.n, new, C.ptr(coarse), C.ptr(parent),
                       ctypes.byref(n_c), C.ptr(self._ws), self._ws.numel(), C.stream_ptr())
            self.sets[new] = CoordSet(coarse[:n_c.value].contiguous(), new)
            self.parent_of[(ts, new)] = parent
        return new

    # -- kernel maps -----------------------------------------------------------------------
    def kernel_map(self, ts_in, ts_out, kernel_size, dilation=1):
        """Map of a conv reading the set at ts_in and writing the set at ts_out."""
        key = (ts_in, ts_out, kernel_size, dilation)
        km = self.kmaps.get(key)
        if km is None:
            cin, cout = self.sets[ts_in].ensure_hash(), self.sets[ts_out]
            K = kernel_size ** 3
            with torch.cuda.device(self.device):
                nbr = torch.empty((K, cout.n), dtype=torch.int32, device=self.device)
                pairs = torch.empty(K, dtype=torch.int32, device=self.device)
                C.call('osb_kernel_map_build', C.ptr(cout.coords), cout.n, C.ptr(cin.slots), cin.cap,
                       kernel_size, kernel_size, kernel_size, ts_in * dilation, C.ptr(nbr), C.ptr(pairs),
                       C.stream_ptr())
            km = KernelMap(nbr, K, cin.n, cout.n, pairs)
            self.kmaps[key] = km
        return km

    def coords_external(self, ts):
        """int32 [N,4] in the caller's row order (stride 1) / internal order (coarser sets)."""
        c = self.sets[ts].coords
        return c[self.inv_perm.long()] if ts == 1 else c
