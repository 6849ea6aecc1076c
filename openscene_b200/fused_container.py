"""A fixed-stride container for fused point features, readable straight into device memory.

The reference stores one ``torch.save({'feat': fp16 [M,C], 'mask_full': bool [N_pts]})`` pickle per scene
(``scripts/feature_fusion/fusion_util.py:86-88``) and ``torch.load``s the whole thing per item
(``dataset/feature_loader.py:96-100``) -- 66-235 GB per dataset, deserialised on the CPU, of which a training item keeps
only the rows of the voxelised points (``feature_loader.py:126-145``, about 40 % at 2 cm).  This container keeps the same
two arrays in a layout a loader can address without unpickling:

    offset 0      header, 64 bytes, little endian:
                  magic 'OSBFEAT1' | int64 n_points | int64 n_rows | int32 channels | int32 dtype code (1 = fp16) |
                  int64 row stride in bytes | int64 offset of the mask bitmap | int64 offset of the rows
    mask bitmap   ceil(n_points / 8) bytes, bit i of byte i//8 (LSB first) = mask_full[i]
    rows          n_rows x row_stride bytes, 4096-byte aligned start, row r = feat[r] (the r-th point with mask_full set)

``FusedFeatureFile.read_remapped`` fuses the loader's remap into the read: the row numbers a voxelised item needs are
computed from the bitmap (a rank query per representative point), ONLY those rows are read from the memory-mapped file into
a pinned staging buffer, in voxel order, and copied to the device asynchronously -- the tensor that arrives IS
``FusedFeatureLoader.__getitem__``'s ``feat_3d`` (and ``mask``), bit for bit.  ``read_all`` + ``remap_fused_features``
(csrc/remap.cu) is the equivalent two-step path for callers that want the whole scene resident.

This module is file plumbing: it owns no kernel and is importable without a GPU (``device`` may be any torch device).
"""
import os
import struct

import numpy as np
import torch

MAGIC = b'OSBFEAT1'
_HEADER = struct.Struct('<8sqqiiqqq')          # 56 bytes used, padded to 64
HEADER_BYTES = 64
_DTYPES = {1: (np.float16, torch.float16)}


def _align(x, a):
    return (x + a - 1) // a * a


def write_container(path, feat, mask_full):
    """feat: [M, C] fp16 (torch or numpy; a trailing singleton axis as in old files is dropped, feature_loader.py:117-118);
    mask_full: bool [N_pts] with exactly M set entries."""
    feat = torch.as_tensor(feat)
    if feat.dim() > 2:
        feat = feat[..., 0]
    feat = feat.to(torch.float16).contiguous().cpu().numpy()
    mask = np.ascontiguousarray(torch.as_tensor(mask_full).bool().cpu().numpy())
    m_rows, c = feat.shape
    if int(mask.sum()) != m_rows:
        raise ValueError(f"write_container: mask_full has {int(mask.sum())} set entries, feat has {m_rows} rows")
    stride = _align(c * 2, 16)
    mask_off = HEADER_BYTES
    bitmap = np.packbits(mask, bitorder='little')
    rows_off = _align(mask_off + bitmap.size, 4096)
    tmp = path + '.tmp'
    with open(tmp, 'wb') as f:
        f.write(_HEADER.pack(MAGIC, mask.size, m_rows, c, 1, stride, mask_off, rows_off).ljust(HEADER_BYTES, b'\0'))
        f.write(bitmap.tobytes())
        f.write(b'\0' * (rows_off - mask_off - bitmap.size))
        if stride == c * 2:
            f.write(feat.tobytes())
        else:
            pad = np.zeros((m_rows, stride // 2), dtype=np.float16)
            pad[:, :c] = feat
            f.write(pad.tobytes())
    os.replace(tmp, path)
    return path


def convert_torch_save(src, dst):
    """One of the reference's ``*.pt`` feature files -> container.  The three-key legacy format ('mask' selecting rows of
    'feat', feature_loader.py:112-115) is resolved here, so the container always has the two-key meaning."""
    d = torch.load(src, map_location='cpu', weights_only=False)
    feat, mask_full = d['feat'], torch.as_tensor(d['mask_full']).bool().clone()
    if feat.dim() > 2:
        feat = feat[..., 0]
    if 'mask' in d:
        legacy = torch.as_tensor(d['mask']).bool()
        feat = feat[legacy]
        mask_full[mask_full.clone()] = legacy
    return write_container(dst, feat, mask_full)


class FusedFeatureFile:
    """Memory-mapped view of one container.  Opening reads the header and the bitmap (N_pts / 8 bytes) only."""

    def __init__(self, path):
        self.path = path
        with open(path, 'rb') as f:
            head = f.read(HEADER_BYTES)
        if len(head) < HEADER_BYTES or head[:8] != MAGIC:
            raise ValueError(f"{path}: not a fused-feature container (bad magic)")
        _, self.n_points, self.n_rows, self.channels, code, self.row_stride, self._mask_off, self._rows_off = \
            _HEADER.unpack(head[:_HEADER.size])
        if code not in _DTYPES:
            raise ValueError(f"{path}: unknown dtype code {code}")
        self._np_dtype, self.dtype = _DTYPES[code]
        size = os.path.getsize(path)
        if self.n_points < 0 or self.n_rows < 0 or self.channels <= 0 or self.row_stride < self.channels * 2 or \
                size < self._rows_off + self.n_rows * self.row_stride:
            raise ValueError(f"{path}: truncated or inconsistent container")
        self._map = np.memmap(path, dtype=np.uint8, mode='c')      # copy-on-write mapping: never written, but torch wants a writable buffer
        nb = (self.n_points + 7) // 8
        self._bitmap = np.asarray(self._map[self._mask_off:self._mask_off + nb])
        self._rows = self._map[self._rows_off:self._rows_off + self.n_rows * self.row_stride].view(self._np_dtype) \
            .reshape(self.n_rows, self.row_stride // 2)
        self._rows_t = torch.from_numpy(self._rows)                 # zero-copy view of the mapping: index_select gathers with all host threads
        self._rank = None
        self._staging = None

    # ------------------------------------------------------------------ host-side index arithmetic
    def mask_full(self):
        """bool [N_pts] (numpy)."""
        return np.unpackbits(self._bitmap, count=self.n_points, bitorder='little').astype(bool)

    def _rank_table(self):
        """rank[i] = number of set bits among points [0, i): the feature row of point i when its bit is set."""
        if self._rank is None:
            m = self.mask_full()
            r = np.zeros(self.n_points + 1, dtype=np.int64)
            np.cumsum(m, out=r[1:])
            self._rank = (m, r)
        return self._rank

    def rows_for(self, vox_ind):
        """vox_ind int64 [N_vox] (the voxeliser's representative points, voxel order).  Returns (rows int64 [n_kept] in voxel
        order, mask_vox bool [N_vox]) -- feature_loader.py:127-145 without touching the feature matrix."""
        vox_ind = np.asarray(vox_ind, dtype=np.int64)
        if vox_ind.size and (vox_ind.min() < 0 or vox_ind.max() >= self.n_points):
            raise ValueError("rows_for: representative point index out of range")
        m, r = self._rank_table()
        mask_vox = m[vox_ind]
        return r[vox_ind[mask_vox]], mask_vox

    # ------------------------------------------------------------------ reads
    def _stage(self, n_rows):
        need = max(1, n_rows) * self.channels
        if self._staging is None or self._staging.numel() < need:
            pin = torch.cuda.is_available()
            self._staging = torch.empty(need, dtype=self.dtype, pin_memory=pin)
        return self._staging[:n_rows * self.channels].view(n_rows, self.channels)

    def read_rows(self, rows, device='cuda', chunk_rows=65536):
        """Feature rows `rows` (any order, repeats allowed) -> [len(rows), C] on `device`.  Rows are gathered from the file
        mapping into a pinned staging buffer chunk by chunk; every chunk's host-to-device copy is asynchronous, so the next
        chunk's page faults overlap it."""
        rows = np.asarray(rows, dtype=np.int64)
        n = rows.size
        dev = torch.device(device)
        out = torch.empty((n, self.channels), dtype=self.dtype, device=dev)
        if n == 0:
            return out
        if rows.min() < 0 or rows.max() >= self.n_rows:
            raise ValueError("read_rows: row index out of range")
        stage = self._stage(n)
        idx = torch.from_numpy(rows)
        src = self._rows_t if self.row_stride == self.channels * 2 else self._rows_t[:, :self.channels]
        for a in range(0, n, chunk_rows):
            b = min(n, a + chunk_rows)
            if src.is_contiguous():
                torch.index_select(src, 0, idx[a:b], out=stage[a:b])
            else:
                stage[a:b] = src[idx[a:b]]
            out[a:b].copy_(stage[a:b], non_blocking=True)
        if dev.type == 'cuda':
            torch.cuda.current_stream(dev).synchronize()           # the staging buffer is reused by the next call
        return out

    def read_all(self, device='cuda'):
        """(feat [M,C], mask_full bool [N_pts]) on `device`: the content of the reference's dict."""
        feat = self.read_rows(np.arange(self.n_rows, dtype=np.int64), device)
        return feat, torch.from_numpy(self.mask_full()).to(device)

    def read_remapped(self, vox_ind, split='train', device='cuda'):
        """What ``FusedFeatureLoader.__getitem__`` hands on after voxelisation (feature_loader.py:101-172), read straight from
        the file: split == 'train' -> (rows of the voxels that have a feature, voxel order; mask per voxel); otherwise ->
        (one row per voxel, zeros where there is none; mask per voxel).  vox_ind: int64 [N_vox], host or device."""
        vi = torch.as_tensor(vox_ind).detach().cpu().numpy()
        rows, mask_vox = self.rows_for(vi)
        dev = torch.device(device)
        kept = self.read_rows(rows, dev)
        mask_t = torch.from_numpy(mask_vox).to(dev)
        if split == 'train':
            return kept, mask_t
        full = torch.zeros((vi.size, self.channels), dtype=self.dtype, device=dev)
        full[mask_t] = kept
        return full, mask_t
