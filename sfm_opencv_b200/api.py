"""Host-side mirror of the reference's interface for the hot path, over the C ABI.

Names and argument meaning follow OpenCV_SFM/NViewReconstuct.cpp:
    match_features(query, train)                 :873-913
    match_features_for_all(descriptor_for_all)   :850-871
    reconstruct(K, R1, T1, R2, T2, p1, p2)       :1117-1159
    ReprojectCost / bundle_adjustment residuals  :142-184, :1187-1211
All compute happens in libsfm_b200.so on the GPU; this module only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import SfmError

RATIO = 0.6          # NViewReconstuct.cpp:884
DIST_FLOOR = 10.0    # :901
GATE_MULT = 5.0      # :901

MATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"),
                        ("distance", "<f4")])      # == cv::DMatch
KNN_DTYPE = np.dtype([("trainIdx0", "<i4"), ("trainIdx1", "<i4"), ("distance0", "<f4"),
                      ("distance1", "<f4")])


class PairLists:
    """Per-pair views of one CSR result array (matches_for_all[i] of the reference): behaves
    like a list of arrays, but a view is only created when a pair is looked at -- building
    19,900 numpy slices eagerly costs more host time than fetching the matches."""

    def __init__(self, flat: np.ndarray, offsets: np.ndarray):
        self.flat, self.offsets = flat, offsets

    def __len__(self):
        return len(self.offsets) - 1

    def __getitem__(self, p):
        if isinstance(p, slice):
            return [self[i] for i in range(*p.indices(len(self)))]
        if p < 0:
            p += len(self)
        if not 0 <= p < len(self):
            raise IndexError(p)
        return self.flat[self.offsets[p]:self.offsets[p + 1]]

    def __iter__(self):
        return (self[p] for p in range(len(self)))


def _ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def _pair_array(pairs) -> np.ndarray:
    """[n,2] int32 view of a pair list.  An int32 ndarray passes through untouched (a caller's
    std::vector<int> pair list costs nothing); a Python list of tuples costs ~0.4 us per pair."""
    if isinstance(pairs, np.ndarray):
        return np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
    return np.asarray(list(pairs), np.int32).reshape(-1, 2)


class Context:
    """One sfm_ctx: one GPU, one host thread."""

    def __init__(self, device: int = 0):
        self._lib = _capi.load()
        err = C.c_int(0)
        self._h = self._lib.sfm_create(device, C.byref(err))
        if not self._h:
            raise SfmError(err.value, self._lib.sfm_last_error(None).decode())
        self.device = device
        self.n_desc: list[int] = []
        self._arenas: dict[str, tuple[int, int]] = {}
        self.last_d2h_bytes = 0

    def close(self):
        if getattr(self, "_h", None):
            for p, _ in self._arenas.values():
                self._lib.sfm_host_free(p)
            self._arenas = {}
            self._lib.sfm_destroy(self._h)
            self._h = None

    def pinned_empty(self, shape, dtype, name: str) -> np.ndarray:
        """A numpy array backed by pinned host memory owned by this context (for callers
        that want asynchronous, full-rate H2D copies of their descriptor banks)."""
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) * dt.itemsize
        return self._pinned(name, max(n, 1))[:n].view(dt).reshape(shape)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise SfmError(rc, self._lib.sfm_last_error(self._h).decode())

    @property
    def launch_count(self) -> int:
        return int(self._lib.sfm_launch_count(self._h))

    @property
    def last_rechecked_rows(self) -> int:
        """Rows of the last match-only call that were recomputed exactly (sfm_last_rechecked_rows)."""
        return int(self._lib.sfm_last_rechecked_rows(self._h))

    # ------------------------------------------------------------------ matching
    def upload_descriptors(self, descriptor_for_all, norm: str = "l2", overlap: bool = False):
        """descriptor_for_all: list of [n_i,128] arrays, float32 (as cv::SIFT gives) or uint8.
        overlap=True queues the transfer and returns (sfm_upload_descriptors_async): the next
        match_pairs overlaps it with the matching kernels and raises its validation errors; the
        arrays must stay alive (and unchanged) until then.
        norm="hamming2": list of [n_i,B<=64] uint8 binary descriptors (AKAZE: B=61), matched
        with cv::NORM_HAMMING2 as the live reference does (NViewReconstuct.cpp:797,876)."""
        descs = [np.ascontiguousarray(d) for d in descriptor_for_all]
        if not descs:
            raise SfmError(_capi.SFM_E_INVALID, "empty image list")
        if norm == "hamming2":
            if any(d.dtype != np.uint8 for d in descs):
                raise SfmError(_capi.SFM_E_INVALID, "binary descriptors must be uint8")
            widths = {d.shape[1] for d in descs if d.ndim == 2}
            width = widths.pop() if len(widths) == 1 else -1
            n = np.array([d.shape[0] for d in descs], np.int32)
            ptrs = (C.c_void_p * len(descs))(*[d.ctypes.data for d in descs])
            self._check(self._lib.sfm_upload_descriptors_bin(self._h, len(descs), ptrs,
                                                             _ptr(n, C.c_int32), width))
            self.n_desc = [int(x) for x in n]
            return
        if norm != "l2":
            raise SfmError(_capi.SFM_E_INVALID, "norm must be 'l2' or 'hamming2'")
        is_u8 = all(d.dtype == np.uint8 for d in descs)
        if not is_u8:
            descs = [np.ascontiguousarray(d, dtype=np.float32) for d in descs]
        dims = {d.shape[1] for d in descs if d.ndim == 2}
        dim = dims.pop() if len(dims) == 1 else -1
        n = np.array([d.shape[0] for d in descs], np.int32)
        ptrs = (C.c_void_p * len(descs))(*[d.ctypes.data for d in descs])
        if overlap:
            self._pending_upload = descs            # keep the host arrays alive
            self._check(self._lib.sfm_upload_descriptors_async(self._h, len(descs), ptrs,
                                                               _ptr(n, C.c_int32), dim, 1 if is_u8 else 4))
        else:
            fn = self._lib.sfm_upload_descriptors_u8 if is_u8 else self._lib.sfm_upload_descriptors
            self._check(fn(self._h, len(descs), ptrs, _ptr(n, C.c_int32), dim))
        self.n_desc = [int(x) for x in n]

    # ------------------------------------------------------------------ sharded upload
    def bank_layout(self, n_desc, overlap: bool = False):
        """Lays out the bank for len(n_desc) images without data (sfm_bank_layout): the same call
        on every GPU; images then arrive by bank_upload_range (from this host) or from a peer
        (all-gather into bank_rows_dev / bank_copy_peer) + bank_commit.
        overlap=True (here and in bank_upload_range / bank_commit): the staged form -- everything
        is queued on the upload stream (upload_stream_torch) and the next match_pairs overlaps
        matching with the arrival of the later ranges (sfm_bank_*_async)."""
        n = np.ascontiguousarray(n_desc, np.int32)
        fn = self._lib.sfm_bank_layout_async if overlap else self._lib.sfm_bank_layout
        self._check(fn(self._h, len(n), _ptr(n, C.c_int32), 128))
        self.n_desc = [int(x) for x in n]
        self._pending_upload = []

    def upload_stream_torch(self):
        """The context's upload stream as a torch.cuda.ExternalStream: collectives issued under
        `with torch.cuda.stream(...)` are ordered with the asynchronous uploads and commits."""
        import torch
        ptr = self._lib.sfm_upload_stream(self._h)
        return torch.cuda.ExternalStream(int(ptr), device=torch.device("cuda", self.device))

    def bank_upload_range(self, first_img: int, descs, overlap: bool = False):
        """Host rows of images first_img .. first_img + len(descs) - 1 -> their bank rows."""
        descs = [np.ascontiguousarray(d) for d in descs]
        if not descs:
            return
        is_u8 = all(d.dtype == np.uint8 for d in descs)
        if not is_u8:
            descs = [np.ascontiguousarray(d, dtype=np.float32) for d in descs]
        for k, d in enumerate(descs):
            if d.ndim != 2 or d.shape != (self.n_desc[first_img + k], 128):
                raise SfmError(_capi.SFM_E_INVALID, "descriptor matrix does not match the bank layout")
        ptrs = (C.c_void_p * len(descs))(*[d.ctypes.data for d in descs])
        if overlap:
            if not isinstance(getattr(self, "_pending_upload", None), list):
                self._pending_upload = []
            self._pending_upload.append(descs)       # keep the host arrays alive
            self._check(self._lib.sfm_bank_upload_range_async(self._h, first_img, len(descs), ptrs,
                                                              1 if is_u8 else 4))
        else:
            self._check(self._lib.sfm_bank_upload_range(self._h, first_img, len(descs), ptrs,
                                                        1 if is_u8 else 4))

    def bank_commit(self, first_img: int, n_img: int, overlap: bool = False):
        fn = self._lib.sfm_bank_commit_async if overlap else self._lib.sfm_bank_commit
        self._check(fn(self._h, first_img, n_img))

    # ------------------------------------------------------------------ peer push exchange
    PEER_HANDLE_BYTES = 160

    def peer_export(self) -> bytes:
        """Handle of this context's bank and mailbox (sfm_peer_export; after bank_layout)."""
        buf = (C.c_uint8 * self.PEER_HANDLE_BYTES)()
        self._check(self._lib.sfm_peer_export(self._h, buf))
        return bytes(buf)

    def peer_connect(self, my_rank: int, handles):
        """handles: the peer_export() bytes of every rank, in rank order (own entry included)."""
        blob = b"".join(handles)
        if len(blob) != self.PEER_HANDLE_BYTES * len(handles):
            raise SfmError(_capi.SFM_E_INVALID, "malformed peer handle list")
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self._check(self._lib.sfm_peer_connect(self._h, my_rank, len(handles), buf))

    def peer_disconnect(self):
        self._check(self._lib.sfm_peer_disconnect(self._h))

    def bank_ready(self, tag: int):
        self._check(self._lib.sfm_bank_ready_async(self._h, tag & 0xFFFFFFFF))

    def bank_push_range(self, first_img: int, n_img: int, slot: int, tag: int):
        self._check(self._lib.sfm_bank_push_range_async(self._h, first_img, n_img, slot, tag & 0xFFFFFFFF))

    def bank_pull_commit(self, src_rank: int, first_img: int, n_img: int, slot: int, tag: int):
        self._check(self._lib.sfm_bank_pull_commit_async(self._h, src_rank, first_img, n_img, slot,
                                                         tag & 0xFFFFFFFF))

    def bank_image_rows(self, img: int):
        """(first bank row, padded row count) of an image; rows of consecutive images are contiguous."""
        r0, n = C.c_int64(0), C.c_int64(0)
        self._check(self._lib.sfm_bank_image_rows(self._h, img, C.byref(r0), C.byref(n)))
        return r0.value, n.value

    def bank_rows_dev(self):
        """(device pointer, rows) of the packed u8 bank [rows][128]."""
        n = C.c_int64(0)
        p = self._lib.sfm_bank_rows_dev(self._h, C.byref(n))
        if not p:
            raise SfmError(_capi.SFM_E_NOT_UPLOADED, "no bank layout")
        return int(p), n.value

    def bank_as_torch(self):
        """The packed bank as a torch uint8 CUDA tensor [rows, 128] aliasing the library's memory
        (for torch.distributed collectives over NVLink; torch is plumbing here, nothing more)."""
        import torch
        ptr, rows = self.bank_rows_dev()

        class _Wrap:
            __cuda_array_interface__ = {"shape": (rows, 128), "typestr": "|u1", "data": (ptr, False),
                                        "version": 3, "strides": None}
        return torch.as_tensor(_Wrap(), device=torch.device("cuda", self.device))

    def bank_copy_peer(self, src: "Context", first_img: int, n_img: int):
        """Packed rows of images [first_img, first_img + n_img) from another context of this
        process (usually another GPU: cudaMemcpyPeerAsync) into this bank, committed."""
        self._check(self._lib.sfm_bank_copy_peer(self._h, src._h, first_img, n_img))

    # ------------------------------------------------------------------ query-row shards
    def match_rows_begin(self, pairs, q_first, q_count, ratio=RATIO):
        """Pass 1 of match_features over query rows [q_first[p], q_first[p] + q_count[p]) of every
        pair: returns this shard's min_dist per pair (reduce with MIN over the shards)."""
        pairs = _pair_array(pairs)
        pq = np.ascontiguousarray(pairs[:, 0])
        pt = np.ascontiguousarray(pairs[:, 1])
        qf = np.ascontiguousarray(q_first, np.int32)
        qc = np.ascontiguousarray(q_count, np.int32)
        if not (len(qf) == len(qc) == len(pq)):
            raise SfmError(_capi.SFM_E_INVALID, "q_first / q_count must have one entry per pair")
        md = np.zeros(max(len(pq), 1), np.float32)
        self._check(self._lib.sfm_match_rows_begin(self._h, _ptr(pq, C.c_int32), _ptr(pt, C.c_int32),
                                                   _ptr(qf, C.c_int32), _ptr(qc, C.c_int32), len(pq),
                                                   ratio, _ptr(md, C.c_float)))
        self._rows_counts = qc.copy()
        return md[:len(pq)]

    def match_rows_finish(self, min_dist, dist_floor=DIST_FLOOR, gate_mult=GATE_MULT, want_knn=False):
        """Pass 2 under the reduced min_dist: returns (per-pair match arrays with image-level
        queryIdx, knn rows per pair or None)."""
        md = np.ascontiguousarray(min_dist, np.float32)
        n_pairs = len(self._rows_counts)
        offsets = np.zeros(n_pairs + 1, np.int64)
        knn = np.zeros(max(int(self._rows_counts.sum()), 1), KNN_DTYPE) if want_knn else None
        self._check(self._lib.sfm_match_rows_finish(self._h, _ptr(md, C.c_float), dist_floor, gate_mult,
                                                    _ptr(offsets, C.c_int64),
                                                    knn.ctypes.data if want_knn else None))
        total = int(offsets[n_pairs])
        out = np.zeros(max(total, 1), MATCH_DTYPE)
        if total:
            self._check(self._lib.sfm_fetch_matches(self._h, out.ctypes.data, total))
        self.last_d2h_bytes = total * MATCH_DTYPE.itemsize + offsets.nbytes
        knn_list = None
        if want_knn:
            knn_list, r = [], 0
            for c in self._rows_counts:
                knn_list.append(knn[r:r + int(c)])
                r += int(c)
        return PairLists(out[:total], offsets), knn_list

    def _pinned(self, name: str, nbytes: int) -> np.ndarray:
        """Grow-only pinned host arena (cudaMallocHost through the C ABI) viewed as uint8."""
        cur = self._arenas.get(name)
        if cur is None or cur[1] < nbytes:
            if cur is not None:
                self._lib.sfm_host_free(cur[0])
            cap = max(int(nbytes * 1.25), 1 << 16)
            p = self._lib.sfm_host_alloc(cap)
            if not p:
                raise SfmError(_capi.SFM_E_NOMEM, "pinned host allocation failed")
            cur = (p, cap)
            self._arenas[name] = cur
        buf = (C.c_uint8 * cur[1]).from_address(cur[0])
        return np.frombuffer(buf, np.uint8, nbytes)

    def match_pairs(self, pairs, ratio=RATIO, dist_floor=DIST_FLOOR, gate_mult=GATE_MULT,
                    want_knn=False, copy=True):
        """pairs: iterable of (query_img, train_img). Returns (list of match arrays per pair,
        min_dist[n_pairs], knn list or None).

        Two C-ABI calls: sfm_match_pairs sizes the result (offsets), sfm_fetch_matches copies
        exactly offsets[n_pairs] matches into a pinned buffer.  With copy=False the returned
        arrays are views of that buffer and are overwritten by the next call."""
        pairs = _pair_array(pairs)
        n_pairs = pairs.shape[0]
        pq = np.ascontiguousarray(pairs[:, 0])
        pt = np.ascontiguousarray(pairs[:, 1])
        if ((pq < 0) | (pq >= len(self.n_desc)) | (pt < 0) | (pt >= len(self.n_desc))).any():
            raise SfmError(_capi.SFM_E_INVALID, "pair index out of range")
        offsets = np.zeros(n_pairs + 1, np.int64)
        min_dist = np.zeros(max(n_pairs, 1), np.float32)
        knn = None
        if want_knn:
            rows = int(sum(self.n_desc[q] for q in pq))
            knn = np.zeros(max(rows, 1), KNN_DTYPE)
        rc = self._lib.sfm_match_pairs(
            self._h, _ptr(pq, C.c_int32), _ptr(pt, C.c_int32), n_pairs, ratio, dist_floor,
            gate_mult, None, 0, _ptr(offsets, C.c_int64),
            knn.ctypes.data if want_knn else None, _ptr(min_dist, C.c_float))
        if rc not in (0, _capi.SFM_E_CAPACITY):
            self._check(rc)
        total = int(offsets[n_pairs])
        raw = self._pinned("matches", max(total, 1) * MATCH_DTYPE.itemsize)
        out = raw.view(MATCH_DTYPE)[:total]          # empty when nothing matched (never a stale record)
        if total:
            self._check(self._lib.sfm_fetch_matches(self._h, out.ctypes.data, total))
        if copy:
            out = out.copy()
        self.last_d2h_bytes = total * MATCH_DTYPE.itemsize + offsets.nbytes + 4 * n_pairs
        matches = PairLists(out, offsets) if n_pairs else []
        knn_list = None
        if want_knn:
            knn_list, r = [], 0
            for q in pq:
                knn_list.append(knn[r:r + self.n_desc[q]])
                r += self.n_desc[q]
        return matches, min_dist[:n_pairs], knn_list

    # ------------------------------------------------------------------ matches -> structure
    def upload_keypoints(self, key_points_for_all):
        """key_points_for_all: per image an [n_i,2] float32 array of cv::KeyPoint::pt."""
        kps = [np.ascontiguousarray(k, np.float32).reshape(-1, 2) for k in key_points_for_all]
        n = np.array([k.shape[0] for k in kps], np.int32)
        ptrs = (C.c_void_p * len(kps))(*[k.ctypes.data for k in kps])
        self._check(self._lib.sfm_upload_keypoints(self._h, len(kps), ptrs, _ptr(n, C.c_int32)))

    def _mask_arg(self, mask):
        if mask is None:
            return None, None
        m = np.ascontiguousarray(np.asarray(mask).reshape(-1), np.uint8)
        return m, _ptr(m, C.c_uint8)

    def get_matched_points(self, pair: int, n_matches: int, mask=None):
        """get_matched_points (+ maskout_points) for pair `pair` of the last match_pairs call,
        gathered on the device.  Returns (p1 [n,2], p2 [n,2]) float32."""
        m, pm = self._mask_arg(mask)
        p1 = np.empty((max(n_matches, 1), 2), np.float32)
        p2 = np.empty((max(n_matches, 1), 2), np.float32)
        n = C.c_int64(0)
        self._check(self._lib.sfm_get_matched_points(self._h, pair, pm, _ptr(p1, C.c_float),
                                                     _ptr(p2, C.c_float), n_matches, C.byref(n)))
        return p1[:n.value], p2[:n.value]

    def reconstruct_pair(self, pair: int, n_matches: int, K, R1, T1, R2, T2, mask=None):
        """reconstruct() over the (masked) matches of pair `pair` of the last match_pairs call
        without bringing the match list to the host.  Returns structure [n,3] float64."""
        m, pm = self._mask_arg(mask)
        a = [np.ascontiguousarray(x, np.float64).reshape(-1) for x in (K, R1, T1, R2, T2)]
        out = np.empty((max(n_matches, 1), 3), np.float64)
        n = C.c_int64(0)
        self._check(self._lib.sfm_reconstruct_pair(
            self._h, pair, *[_ptr(x, C.c_double) for x in a], pm, _ptr(out, C.c_double), n_matches,
            C.byref(n)))
        return out[:n.value]

    def match_pairs_resident(self, pairs, ratio=RATIO, dist_floor=DIST_FLOOR,
                             gate_mult=GATE_MULT):
        """Device-resident timing hook: returns (total_matches, knn_kernel_ms, total_ms)."""
        pairs = _pair_array(pairs)
        pq = np.ascontiguousarray(pairs[:, 0])
        pt = np.ascontiguousarray(pairs[:, 1])
        total = C.c_int64(0)
        kms, tms = C.c_float(0), C.c_float(0)
        self._check(self._lib.sfm_match_pairs_resident(
            self._h, _ptr(pq, C.c_int32), _ptr(pt, C.c_int32), pairs.shape[0], ratio, dist_floor,
            gate_mult, C.byref(total), C.byref(kms), C.byref(tms)))
        return total.value, kms.value, tms.value

    # ------------------------------------------------------------------ triangulation
    def triangulate_batch(self, P, xy, want_X4=True, want_xyz=True, iters=0, out_X4=None, out_xyz=None):
        """P: [V,3,4] float32; xy: [V,N,2] float32. Returns (X4 [4,N] f32, xyz [N,3] f64[, ms]).
        out_X4 / out_xyz: caller-owned result arrays (e.g. pinned: pinned_empty) to fill."""
        P = np.ascontiguousarray(P, np.float32).reshape(-1, 3, 4)
        xy = np.ascontiguousarray(xy, np.float32)
        V = P.shape[0]
        if xy.ndim != 3 or xy.shape[0] != V or xy.shape[2] != 2:
            raise SfmError(_capi.SFM_E_INVALID, "xy must be [V,N,2]")
        N = xy.shape[1]
        for o, shp, dt in ((out_X4, (4, N), np.float32), (out_xyz, (N, 3), np.float64)):
            if o is not None and (o.shape != shp or o.dtype != dt or not o.flags.c_contiguous):
                raise SfmError(_capi.SFM_E_INVALID, "output array has the wrong shape / dtype")
        X4 = (out_X4 if out_X4 is not None else np.empty((4, N), np.float32)) if want_X4 else None
        xyz = (out_xyz if out_xyz is not None else np.empty((N, 3), np.float64)) if want_xyz else None
        pX4 = _ptr(X4, C.c_float) if want_X4 and N else None
        pxyz = _ptr(xyz, C.c_double) if want_xyz and N else None
        if iters > 0:
            ms = C.c_float(0)
            self._check(self._lib.sfm_triangulate_batch_timed(
                self._h, _ptr(P, C.c_float), _ptr(xy, C.c_float), V, N, pX4, pxyz, iters,
                C.byref(ms)))
            return X4, xyz, ms.value
        self._check(self._lib.sfm_triangulate_batch(
            self._h, _ptr(P, C.c_float), _ptr(xy, C.c_float), V, N, pX4, pxyz))
        return X4, xyz

    # ------------------------------------------------------------------ residuals
    def reproject_residuals(self, intr, ext, pts, cam_idx, pt_idx, obs_xy, huber_delta=4.0,
                            want_resid=True, want_cost=True, iters=0, out_resid=None):
        intr = np.ascontiguousarray(intr, np.float64).reshape(4)
        ext = np.ascontiguousarray(ext, np.float64).reshape(-1, 6)
        pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
        cam_idx = np.ascontiguousarray(cam_idx, np.int32)
        pt_idx = np.ascontiguousarray(pt_idx, np.int32)
        obs_xy = np.ascontiguousarray(obs_xy, np.float32).reshape(-1, 2)
        n_obs = cam_idx.shape[0]
        if out_resid is not None and (out_resid.shape != (n_obs, 2) or out_resid.dtype != np.float64):
            raise SfmError(_capi.SFM_E_INVALID, "out_resid must be [n_obs,2] float64")
        resid = (out_resid if out_resid is not None else np.empty((n_obs, 2), np.float64)) if want_resid else None
        cost = C.c_double(0)
        args = [self._h, _ptr(intr, C.c_double), _ptr(ext, C.c_double), ext.shape[0],
                _ptr(pts, C.c_double), pts.shape[0], _ptr(cam_idx, C.c_int32),
                _ptr(pt_idx, C.c_int32), _ptr(obs_xy, C.c_float), n_obs, float(huber_delta),
                _ptr(resid, C.c_double) if want_resid and n_obs else None,
                C.byref(cost) if want_cost else None]
        if iters > 0:
            ms = C.c_float(0)
            self._check(self._lib.sfm_reproject_residuals_timed(*args, iters, C.byref(ms)))
            return resid, (cost.value if want_cost else None), ms.value
        self._check(self._lib.sfm_reproject_residuals(*args))
        return resid, (cost.value if want_cost else None)

    def reproject_jacobians(self, intr, ext, pts, cam_idx, pt_idx, obs_xy, want_resid=True,
                            iters=0):
        """Residuals and their Jacobians (what Ceres' autodiff derives from ReprojectCost,
        NViewReconstuct.cpp:1202): returns (resid [n,2] or None, J [n,2,13]) with the columns
        ordered (fx, fy, cx, cy | angle-axis(3), t(3) | X, Y, Z); with iters > 0 also the mean
        kernel time in ms."""
        intr = np.ascontiguousarray(intr, np.float64).reshape(4)
        ext = np.ascontiguousarray(ext, np.float64).reshape(-1, 6)
        pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
        cam_idx = np.ascontiguousarray(cam_idx, np.int32)
        pt_idx = np.ascontiguousarray(pt_idx, np.int32)
        obs_xy = np.ascontiguousarray(obs_xy, np.float32).reshape(-1, 2)
        n_obs = cam_idx.shape[0]
        resid = np.empty((n_obs, 2), np.float64) if want_resid else None
        jac = np.empty((n_obs, 2, 13), np.float64)
        ms = C.c_float(0)
        self._check(self._lib.sfm_reproject_jacobians(
            self._h, _ptr(intr, C.c_double), _ptr(ext, C.c_double), ext.shape[0],
            _ptr(pts, C.c_double), pts.shape[0], _ptr(cam_idx, C.c_int32), _ptr(pt_idx, C.c_int32),
            _ptr(obs_xy, C.c_float), n_obs, _ptr(resid, C.c_double) if want_resid and n_obs else None,
            _ptr(jac, C.c_double) if n_obs else None, iters, C.byref(ms)))
        return (resid, jac, ms.value) if iters > 0 else (resid, jac)

    def estimate_normals(self, pts3d, K: int = 10):
        """estimate_normals(pts3d, K, normals), NViewReconstuct.cpp:551-599 (K = 10 at :1502):
        returns normals [N,3] float64."""
        pts = np.ascontiguousarray(pts3d, np.float64).reshape(-1, 3)
        out = np.empty_like(pts)
        self._check(self._lib.sfm_estimate_normals(self._h, _ptr(pts, C.c_double), pts.shape[0], K,
                                                   _ptr(out, C.c_double)))
        return out

    # ------------------------------------------------------------------ timing hooks
    def timer_start(self):
        self._check(self._lib.sfm_timer_start(self._h))

    def timer_stop(self) -> float:
        """Device time in ms (CUDA events on the context's stream) since timer_start()."""
        ms = C.c_float(0)
        self._check(self._lib.sfm_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def sync(self):
        self._check(self._lib.sfm_sync(self._h))

    def probe_fp64_peak(self, iters: int = 4000) -> float:
        t = C.c_double(0)
        self._check(self._lib.sfm_probe_fp64_peak(self._h, iters, C.byref(t)))
        return t.value

    def probe_i8_peak(self, iters: int = 2000) -> float:
        tops = C.c_double(0)
        self._check(self._lib.sfm_probe_i8_peak(self._h, iters, C.byref(tops)))
        return tops.value


class BAProblem:
    """The residual blocks of one bundle_adjustment() call (NViewReconstuct.cpp:1187-1211), resident
    on the GPU: observation tables are uploaded and range-checked once; evaluate() moves only the
    cameras and points (what changes between LM iterations)."""

    def __init__(self, ctx: Context, n_cam: int, n_pts: int, cam_idx, pt_idx, obs_xy):
        self.ctx = ctx
        cam_idx = np.ascontiguousarray(cam_idx, np.int32)
        pt_idx = np.ascontiguousarray(pt_idx, np.int32)
        obs_xy = np.ascontiguousarray(obs_xy, np.float32).reshape(-1, 2)
        self.n_cam, self.n_pts, self.n_obs = int(n_cam), int(n_pts), int(cam_idx.shape[0])
        h = C.c_void_p(None)
        ctx._check(ctx._lib.sfm_ba_create(ctx._h, self.n_cam, self.n_pts, _ptr(cam_idx, C.c_int32),
                                          _ptr(pt_idx, C.c_int32), _ptr(obs_xy, C.c_float), self.n_obs,
                                          C.byref(h)))
        self._h = h
        self.kernel_ms = 0.0

    def evaluate(self, intr, ext=None, pts=None, huber_delta=4.0, want_resid=True, want_jac=False,
                 want_cost=True):
        """Returns (resid [n_obs,2] | None, jac [n_obs,2,13] | None, cost | None); ext / pts None
        keeps the values of the previous evaluation."""
        intr = np.ascontiguousarray(intr, np.float64).reshape(4)
        pe = pp = None
        if ext is not None:
            ext = np.ascontiguousarray(ext, np.float64).reshape(self.n_cam, 6)
            pe = _ptr(ext, C.c_double)
        if pts is not None:
            pts = np.ascontiguousarray(pts, np.float64).reshape(self.n_pts, 3)
            pp = _ptr(pts, C.c_double)
        resid = np.empty((self.n_obs, 2), np.float64) if want_resid else None
        jac = np.empty((self.n_obs, 2, 13), np.float64) if want_jac else None
        cost, ms = C.c_double(0), C.c_float(0)
        self.ctx._check(self.ctx._lib.sfm_ba_evaluate(
            self.ctx._h, self._h, _ptr(intr, C.c_double), pe, pp, float(huber_delta),
            _ptr(resid, C.c_double) if want_resid else None, _ptr(jac, C.c_double) if want_jac else None,
            C.byref(cost) if want_cost else None, C.byref(ms)))
        self.kernel_ms = ms.value
        return resid, jac, (cost.value if want_cost else None)

    def close(self):
        if getattr(self, "_h", None) and self.ctx._h:
            self.ctx._lib.sfm_ba_destroy(self.ctx._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------- reference-shaped API

def match_features(ctx: Context, query, train, norm: str = "l2", **kw):
    """match_features(query, train, matches), NViewReconstuct.cpp:873: returns the DMatch array.
    norm="hamming2" is the live file's BFMatcher(NORM_HAMMING2) (:876), "l2" the SIFT form."""
    ctx.upload_descriptors([query, train], norm=norm)
    m, _, _ = ctx.match_pairs([(0, 1)], **kw)
    return m[0]


def match_features_for_all(ctx: Context, descriptor_for_all, norm: str = "l2", **kw):
    """match_features_for_all, NViewReconstuct.cpp:850-871: consecutive pairs (i, i+1)."""
    ctx.upload_descriptors(descriptor_for_all, norm=norm)
    pairs = [(i, i + 1) for i in range(len(descriptor_for_all) - 1)]
    m, _, _ = ctx.match_pairs(pairs, **kw)
    return m


def build_projection(K, R, T) -> np.ndarray:
    """proj = fK * [R|T], NViewReconstuct.cpp:1129-1143 (host glue), evaluated as cv::gemm does
    for CV_32F operands so that P is bit-identical to the reference's."""
    RT = np.empty((3, 4), np.float32)
    RT[:, :3] = np.asarray(R, np.float64).astype(np.float32)
    RT[:, 3] = np.asarray(T, np.float64).reshape(3).astype(np.float32)
    fK = np.asarray(K, np.float64).astype(np.float32)
    # cv::gemm's small-matrix path: ((a0*b0 + a1*b1) + a2*b2) in float32, no FMA
    p = fK[:, :, None] * RT[None, :, :]                  # float32 products [r, k, c]
    return ((p[:, 0, :] + p[:, 1, :]) + p[:, 2, :]).astype(np.float32)


def enumerate_observations(inds_2d_to_3d, keypoints_xy):
    """Residual-block order of bundle_adjustment(), NViewReconstuct.cpp:1187-1211 (host glue):
    for img in 0..n-1, for kp in order: if inds_2d_to_3d[img][kp] >= 0 -> one observation
    (camera img, point id, kp.pt).  Returns (cam_idx i32, pt_idx i32, obs_xy f32[n,2])."""
    cam, pt, obs = [], [], []
    for img, (ids, kps) in enumerate(zip(inds_2d_to_3d, keypoints_xy)):
        ids = np.asarray(ids).reshape(-1)
        sel = np.nonzero(ids >= 0)[0]
        cam.append(np.full(sel.size, img, np.int32))
        pt.append(ids[sel].astype(np.int32))
        obs.append(np.asarray(kps, np.float32).reshape(-1, 2)[sel])
    if not cam:
        return np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 2), np.float32)
    return np.concatenate(cam), np.concatenate(pt), np.concatenate(obs).astype(np.float32)


def bundle_adjustment_residuals(ctx: Context, intrinsic, extrinsics, inds_2d_to_3d, keypoints_xy,
                                pts3d, huber_delta=4.0):
    """Evaluates every ReprojectCost block that bundle_adjustment() (NViewReconstuct.cpp:1162-1244)
    hands to Ceres, in its order: returns (residuals [n_obs,2], cost = 0.5*sum rho(|r|^2) with
    HuberLoss(huber_delta) as at :1184, rmse = sqrt(cost / n_obs) as printed at :1237-1238)."""
    cam, pt, obs = enumerate_observations(inds_2d_to_3d, keypoints_xy)
    r, cost = ctx.reproject_residuals(intrinsic, extrinsics, pts3d, cam, pt, obs,
                                      huber_delta=huber_delta)
    n = max(len(cam), 1)
    return r, cost, float(np.sqrt(cost / n))


def save_structure(file_name, rotations, motions, structure, colors):
    """save_structure(file_name, rotations, motions, structure, colors), NViewReconstuct.cpp:
    186-227: the cv::FileStorage YAML the viewer loads, written by the library's own emitter
    (byte-identical to OpenCV's).  rotations: n x 3x3, motions: n x 3x1, structure: [N,3]
    float64, colors: [M,3] uint8 in the reference's b,g,r order."""
    lib = _capi.load()
    R = np.ascontiguousarray(np.asarray(rotations, np.float64).reshape(-1, 9))
    T = np.ascontiguousarray(np.asarray(motions, np.float64).reshape(-1, 3))
    if R.shape[0] != T.shape[0]:
        raise SfmError(_capi.SFM_E_INVALID, "rotations and motions differ in length")
    X = np.ascontiguousarray(np.asarray(structure, np.float64).reshape(-1, 3))
    c = np.ascontiguousarray(np.asarray(colors, np.uint8).reshape(-1, 3))
    rc = lib.sfm_save_structure(str(file_name).encode(), R.shape[0], _ptr(R, C.c_double),
                                _ptr(T, C.c_double), X.shape[0], _ptr(X, C.c_double),
                                c.shape[0], _ptr(c, C.c_uint8))
    if rc:
        raise SfmError(rc, f"cannot write {file_name}")


def write_ply_binary(path, xyz, normals, rgb, crlf: bool = True):
    """write_ply_binary(path, points), NViewReconstuct.cpp:229-294: 27 bytes per vertex
    (x y z nx ny nz float32, r g b uint8), NaN vertices skipped.  crlf=True gives the header
    line ends the reference produces on its platform (bundled Viewer/structure_ba.ply)."""
    lib = _capi.load()
    v = np.ascontiguousarray(np.concatenate([np.asarray(xyz, np.float32).reshape(-1, 3),
                                             np.asarray(normals, np.float32).reshape(-1, 3)], 1))
    c = np.ascontiguousarray(np.asarray(rgb, np.uint8).reshape(-1, 3))
    if c.shape[0] != v.shape[0]:
        raise SfmError(_capi.SFM_E_INVALID, "points and colours differ in length")
    rc = lib.sfm_write_ply_binary(str(path).encode(), v.shape[0], _ptr(v, C.c_float),
                                  _ptr(c, C.c_uint8), int(crlf))
    if rc:
        raise SfmError(rc, f"cannot write {path}")


def reconstruct(ctx: Context, K, R1, T1, R2, T2, p1, p2):
    """reconstruct(K,R1,T1,R2,T2,p1,p2,structure), NViewReconstuct.cpp:1117: returns structure
    [N,3] float64 (Point3d); raises on empty input where the reference returns -1."""
    p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 2)
    p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 2)
    if p1.shape[0] == 0 or p2.shape[0] == 0:
        raise SfmError(_capi.SFM_E_INVALID, "[Err]: empty 2d points.")
    P = np.stack([build_projection(K, R1, T1), build_projection(K, R2, T2)])
    _, xyz = ctx.triangulate_batch(P, np.stack([p1, p2]), want_X4=False)
    return xyz
