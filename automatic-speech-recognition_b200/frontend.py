"""Host-side driver of the CUDA front-end: configuration, table upload, packing of
variable-length utterances, and the batch call through the C-ABI.

Mirrors what the loop body of ``process_audios`` does per file
(/root/reference/preprocess.py:67-89), for a whole batch at once."""
import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib, tables

PCM_ALIGN_INT16 = 8     # samples: per-utterance offsets are 16-byte aligned
PCM_ALIGN_F32 = 4
OUT_ALIGN = 4           # floats


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def _device_ptr(x):
    """Raw device pointer of a torch CUDA tensor / __cuda_array_interface__ object, else None."""
    if hasattr(x, "data_ptr") and hasattr(x, "is_cuda"):
        return int(x.data_ptr()) if x.is_cuda else None
    if hasattr(x, "__cuda_array_interface__"):
        return int(x.__cuda_array_interface__["data"][0])
    return None


@dataclass
class FrontendConfig:
    """Same knobs the reference reads from ``args`` (preprocess.py:59-63) plus the
    speechpy defaults it relies on implicitly and the switches SURVEY.md asks for."""
    sample_rate: int = 16000
    frame_length: float = 25.0          # ms  (args.frame_length, las/arguments.py:33-36)
    frame_step: float = 10.0            # ms  (args.frame_step,   las/arguments.py:37-40)
    feat_dim: int = 13                  # args.feat_dim
    feat_type: str = "mfcc"             # args.feat_type: 'mfcc' | 'fbank'
    cmvn: bool = True                   # args.cmvn
    num_filters: int = 40               # speechpy mfcc default; ignored for fbank (= feat_dim)
    fft_length: int = 512
    low_frequency: float = 0            # speechpy turns 0 into 300 Hz
    high_frequency: Optional[float] = None
    dc_elimination: bool = True
    delta_mode: str = "speechpy_as_shipped"     # | 'time_regression'
    bin_map: str = "coefficients_plus_one"      # | 'nfft_plus_one'
    fbank_log: bool = False             # reference 'fbank' = linear mel energies (mfe)
    window: Optional[np.ndarray] = None  # None = rectangular (reference)
    preemph: float = 0.0                # 0 = off (reference never calls preemphasis)
    pcm_dtype: str = "int16"            # | 'float32'
    speeds: Tuple[float, ...] = (0.9, 1.1)      # preprocess.py:160

    @classmethod
    def from_args(cls, args, **overrides):
        kw = dict(frame_length=args.frame_length, frame_step=args.frame_step,
                  feat_dim=args.feat_dim, feat_type=args.feat_type, cmvn=bool(args.cmvn))
        if getattr(args, "sample_rate", None):
            kw["sample_rate"] = args.sample_rate
        kw.update(overrides)
        return cls(**kw)

    # speechpy.processing.stack_frames rounding rules
    @property
    def frame_len(self):
        return int(np.round(self.sample_rate * self.frame_length / 1000.0))

    @property
    def hop(self):
        return int(np.round(self.sample_rate * self.frame_step / 1000.0))

    @property
    def n_filters(self):
        return self.feat_dim if self.feat_type == "fbank" else self.num_filters

    @property
    def out_width(self):
        return self.feat_dim * (3 if self.cmvn else 1)


def make_fe_config(c: "FrontendConfig"):
    """FrontendConfig -> (ctypes fe_config, keep-alive dict of the numpy tables it points
    to, {speed: index}).  Pure host code (no CUDA): also used by the CPU replay of the
    kernel dataflow in tests/host_sim."""
    if c.feat_type not in ("mfcc", "fbank"):
        raise ValueError("feat_type must be 'mfcc' or 'fbank'")
    if c.fft_length != tables.NFFT:
        raise ValueError("only fft_length=512 is supported")
    nf = c.n_filters
    fb = tables.mel_filterbank_dense(nf, c.sample_rate, c.low_frequency, c.high_frequency, c.bin_map)
    row_start, first_bin, w = tables.filterbank_csr(fb)
    keep = {
        "row_start": np.ascontiguousarray(row_start, dtype=np.int32),
        "first_bin": np.ascontiguousarray(first_bin, dtype=np.int32),
        "w": np.ascontiguousarray(w, dtype=np.float32),
        "tw256": np.ascontiguousarray(tables.twiddles_256().reshape(-1), dtype=np.float32),
        "tw512": np.ascontiguousarray(tables.twiddles_512().reshape(-1), dtype=np.float32),
    }
    cfg = _lib.FeConfig()
    cfg.abi_version = _lib.FE_ABI_VERSION
    cfg.sample_rate = int(c.sample_rate)
    cfg.frame_len, cfg.hop, cfg.nfft = c.frame_len, c.hop, c.fft_length
    cfg.num_filters, cfg.feat_dim = nf, int(c.feat_dim)
    cfg.feat_type = _lib.FE_FEAT_MFCC if c.feat_type == "mfcc" else _lib.FE_FEAT_FBANK
    cfg.cmvn = int(bool(c.cmvn))
    cfg.delta_mode = {"speechpy_as_shipped": _lib.FE_DELTA_SPEECHPY,
                      "time_regression": _lib.FE_DELTA_TIME_REGRESSION}[c.delta_mode]
    cfg.fbank_log = int(bool(c.fbank_log))
    cfg.dc_elimination = int(bool(c.dc_elimination))
    cfg.pcm_dtype = {"int16": _lib.FE_PCM_INT16, "float32": _lib.FE_PCM_FLOAT32}[c.pcm_dtype]
    cfg.preemph = float(c.preemph or 0.0)
    cfg.fb_nnz = int(keep["w"].size)
    cfg.fb_row_start = _ptr(keep["row_start"], C.c_int32)
    cfg.fb_first_bin = _ptr(keep["first_bin"], C.c_int32)
    cfg.fb_weights = _ptr(keep["w"], C.c_float)
    if c.feat_type == "mfcc":
        keep["dct"] = np.ascontiguousarray(tables.dct_ortho(nf, c.feat_dim).reshape(-1), dtype=np.float32)
        cfg.dct = _ptr(keep["dct"], C.c_float)
    if c.window is not None:
        win = np.ascontiguousarray(c.window, dtype=np.float32)
        if win.shape != (c.frame_len,):
            raise ValueError("window must have frame_len entries")
        keep["window"] = win
        cfg.window = _ptr(win, C.c_float)
    cfg.tw256 = _ptr(keep["tw256"], C.c_float)
    cfg.tw512 = _ptr(keep["tw512"], C.c_float)
    speeds = [s for s in c.speeds if abs(s - 1.0) > 1e-12]
    speed_index = {}
    if speeds:
        ups, downs, taps = [], [], []
        for i, s in enumerate(speeds):
            up, down = tables.speed_ratio(s)
            ups.append(up); downs.append(down)
            taps.append(tables.resampler_taps(s).astype(np.float32).reshape(-1))
            speed_index[round(float(s), 6)] = i
        keep["sp_up"] = np.asarray(ups, dtype=np.int32)
        keep["sp_down"] = np.asarray(downs, dtype=np.int32)
        keep["sp_taps"] = np.ascontiguousarray(np.concatenate(taps), dtype=np.float32)
        cfg.n_speeds = len(speeds)
        cfg.speed_up = _ptr(keep["sp_up"], C.c_int32)
        cfg.speed_down = _ptr(keep["sp_down"], C.c_int32)
        cfg.speed_taps = _ptr(keep["sp_taps"], C.c_float)
        cfg.speed_ntaps = tables.RESAMPLE_TAPS
    return cfg, keep, speed_index


def num_frames(n_samples, frame_len=400, hop=160):
    """floor((N - frame_len) / hop), clamped at 0 (speechpy stack_frames, zero_padding=False)."""
    return int(_lib.load().fe_num_frames(int(n_samples), int(frame_len), int(hop)))


def pack_pcm(pcm_list: Sequence[np.ndarray], dtype=np.int16):
    """Concatenate utterances into one buffer with 16-byte aligned starts.
    Returns (packed, offsets[n] int64, lengths[n] int64)."""
    align = PCM_ALIGN_INT16 if np.dtype(dtype) == np.int16 else PCM_ALIGN_F32
    lengths = np.fromiter((len(p) for p in pcm_list), dtype=np.int64, count=len(pcm_list))
    padded = (lengths + align - 1) // align * align
    offsets = np.zeros(len(pcm_list), dtype=np.int64)
    if len(pcm_list) > 1:
        np.cumsum(padded[:-1], out=offsets[1:])
    total = int(padded.sum())
    packed = np.zeros(max(total, align), dtype=dtype)
    for p, o, n in zip(pcm_list, offsets, lengths):
        packed[o:o + n] = p
    return packed, offsets, lengths


class Frontend:
    """One handle = one (host thread, GPU).  Not thread-safe (same as the C-ABI)."""

    def __init__(self, config: Optional[FrontendConfig] = None, device: int = 0):
        self.config = config or FrontendConfig()
        self.device = device
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self._pinned_ptrs = []
        rc = self._lib.fe_create(int(device), C.byref(self._h))
        if rc != 0:
            msg = self._lib.fe_last_error(None)
            raise RuntimeError("fe_create(device=%d) failed (%d): %s" % (device, rc, msg.decode() if msg else ""))
        self._configure()

    # -- pinned host staging ---------------------------------------------------
    def pinned(self, n, dtype=np.uint8):
        """A page-locked numpy array of n elements owned by this handle (freed in close())."""
        dt = np.dtype(dtype)
        ptr = C.c_void_p()
        self._check(self._lib.fe_host_alloc(self._h, int(max(n, 1)) * dt.itemsize, C.byref(ptr)), "fe_host_alloc")
        self._pinned_ptrs.append(ptr.value)
        raw = (C.c_uint8 * (int(max(n, 1)) * dt.itemsize)).from_address(ptr.value)
        return np.frombuffer(raw, dtype=dt, count=int(max(n, 1)))

    def free_pinned(self, arrays):
        """Release page-locked arrays obtained from ``pinned`` (a staging buffer that has been replaced by a larger one)."""
        for a in arrays:
            ptr = a.__array_interface__["data"][0]
            if ptr in self._pinned_ptrs:
                self._pinned_ptrs.remove(ptr)
                self._check(self._lib.fe_host_free(self._h, C.c_void_p(ptr)), "fe_host_free")

    # -- lifecycle ---------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            for p in getattr(self, "_pinned_ptrs", []):
                self._lib.fe_host_free(self._h, C.c_void_p(p))
            self._pinned_ptrs = []
            self._lib.fe_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.fe_last_error(self._h)
            raise RuntimeError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))

    # -- configuration -----------------------------------------------------
    def _configure(self):
        cfg, self._keep, self._speed_index = make_fe_config(self.config)
        self._check(self._lib.fe_configure(self._h, C.byref(cfg)), "fe_configure")

    # -- helpers -----------------------------------------------------------
    def speed_indices(self, speeds):
        """Per-utterance speed values -> int32 indices into the configured table (-1 = 1.0)."""
        if speeds is None:
            return None
        idx = np.full(len(speeds), -1, dtype=np.int32)
        for i, s in enumerate(speeds):
            if s is None or abs(float(s) - 1.0) < 1e-12:
                continue
            key = round(float(s), 6)
            if key not in self._speed_index:
                raise ValueError("speed %r not in FrontendConfig.speeds %r" % (s, self.config.speeds))
            idx[i] = self._speed_index[key]
        return idx

    def plan(self, lengths, speed_idx=None):
        """Host-only: (out_offsets[n+1] in floats, n_frames[n]) for a batch."""
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        n = lengths.size
        out_off = np.zeros(n + 1, dtype=np.int64)
        nfr = np.zeros(max(n, 1), dtype=np.int32)
        sp = None if speed_idx is None else np.ascontiguousarray(speed_idx, dtype=np.int32)
        self._check(self._lib.fe_plan(self._h, _ptr(lengths, C.c_int64), n,
                                      None if sp is None else _ptr(sp, C.c_int32),
                                      _ptr(out_off, C.c_int64), _ptr(nfr, C.c_int32)), "fe_plan")
        return out_off, nfr[:n]

    def run_packed(self, pcm, offsets, lengths, speed_idx=None, gain=None, out=None, stream=None):
        """The batch hot path.  ``pcm``/``out`` are numpy arrays (host) or CUDA tensors
        (device, stays resident).  Returns (out, out_offsets, n_frames)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        n = lengths.size
        sp = None if speed_idx is None else np.ascontiguousarray(speed_idx, dtype=np.int32)
        gn = None if gain is None else np.ascontiguousarray(gain, dtype=np.float32)
        out_off = np.zeros(n + 1, dtype=np.int64)
        nfr = np.zeros(max(n, 1), dtype=np.int32)
        pcm_dev = _device_ptr(pcm)
        if pcm_dev is None:
            want = np.int16 if self.config.pcm_dtype == "int16" else np.float32
            pcm = np.ascontiguousarray(pcm, dtype=want)
            pcm_ptr = pcm.ctypes.data
        else:
            pcm_ptr = pcm_dev
        if out is None:
            plan_off, _ = self.plan(lengths, sp)
            total = int(plan_off[-1])
            if pcm_dev is not None:
                import torch
                out = torch.empty(max(total, 1), dtype=torch.float32, device=pcm.device)
            else:
                out = np.empty(max(total, 1), dtype=np.float32)
        out_dev = _device_ptr(out)
        if out_dev is None:
            if not (isinstance(out, np.ndarray) and out.dtype == np.float32 and out.flags.c_contiguous):
                raise ValueError("out must be a C-contiguous float32 array")
            out_ptr, cap = out.ctypes.data, out.size
        else:
            out_ptr, cap = out_dev, int(out.numel()) if hasattr(out, "numel") else int(np.prod(out.shape))
        self._check(self._lib.fe_run(
            self._h, C.c_void_p(pcm_ptr), _ptr(offsets, C.c_int64), _ptr(lengths, C.c_int64), n,
            None if sp is None else _ptr(sp, C.c_int32), None if gn is None else _ptr(gn, C.c_float),
            C.c_void_p(out_ptr), cap, _ptr(out_off, C.c_int64), _ptr(nfr, C.c_int32),
            C.c_void_p(int(stream)) if stream else None), "fe_run")
        return out, out_off, nfr[:n]

    def decode_flac(self, buf, files, n_files, total_bytes, pcm_total, pcm=None, stream=None):
        """FLAC bytes -> int16 PCM on the GPU (``fe_decode_flac``).  ``buf``: numpy uint8 (host) or a CUDA
        uint8 tensor padded by 4096 bytes; ``files``: the fe_flac_file array ``audio_io.load_flac_batch``
        planned.  ``pcm``: optional destination (CUDA int16 tensor or numpy int16); by default a CUDA tensor
        is allocated, so the samples stay in HBM for ``run_packed``.  Returns pcm."""
        dev = _device_ptr(buf)
        in_ptr = dev if dev is not None else buf.ctypes.data
        if pcm is None:
            import torch
            pcm = torch.zeros(max(int(pcm_total), 8), dtype=torch.int16, device="cuda:%d" % self.device)
            torch.cuda.current_stream(self.device).synchronize()      # the handle's stream does not wait for torch's
        pdev = _device_ptr(pcm)
        out_ptr = pdev if pdev is not None else pcm.ctypes.data
        cap = int(pcm.numel()) if hasattr(pcm, "numel") else int(pcm.size)
        status = np.zeros(max(n_files, 1), dtype=np.int32)
        rc = self._lib.fe_decode_flac(self._h, C.c_void_p(in_ptr), int(total_bytes), files, int(n_files),
                                      C.c_void_p(out_ptr), cap, _ptr(status, C.c_int32),
                                      C.c_void_p(int(stream)) if stream else None)
        self.flac_status = status[:n_files]
        self._check(rc, "fe_decode_flac")
        return pcm

    def flac_ms(self):
        ms = (C.c_float * 3)()
        self._check(self._lib.fe_get_flac_ms(self._h, ms), "fe_get_flac_ms")
        return {"scan": ms[0], "decode": ms[1], "validate": ms[2]}

    def split(self, out, out_offsets, n_frames, copy=False):
        """Flat output -> list of per-utterance arrays, (L, D, 3) with cmvn else (L, D)."""
        if _device_ptr(out) is not None:
            self.sync()                    # fe_run on device buffers is asynchronous on the handle's stream
            out = out.cpu().numpy()
        c = self.config
        res = []
        for i, L in enumerate(n_frames):
            o = int(out_offsets[i])
            a = out[o:o + int(L) * c.out_width]
            a = a.reshape((int(L), c.feat_dim, 3) if c.cmvn else (int(L), c.feat_dim))
            res.append(a.copy() if copy else a)
        return res

    def extract(self, pcm_list, speeds=None, gains=None, copy=False):
        """List of 1-D PCM arrays -> list of float32 feature arrays (host in, host out)."""
        if len(pcm_list) == 0:
            return []
        dtype = np.int16 if self.config.pcm_dtype == "int16" else np.float32
        packed, off, lens = pack_pcm(pcm_list, dtype)
        out, out_off, nfr = self.run_packed(packed, off, lens, self.speed_indices(speeds), gains)
        return self.split(out, out_off, nfr, copy=copy)

    def perturb_packed(self, packed, off, lens, speeds=None, gains=None):
        """Speed / volume perturbation of a packed int16 batch (host in, host out).
        Returns (dst, dst_offsets[n + 1], dst_lengths[n]); utterance i is
        dst[dst_offsets[i] : dst_offsets[i] + dst_lengths[i]], starts 16-byte aligned."""
        packed = np.ascontiguousarray(packed, dtype=np.int16)
        off = np.ascontiguousarray(off, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int64)
        n = lens.size
        sp = self.speed_indices(speeds)
        gn = None if gains is None else np.ascontiguousarray(gains, dtype=np.float32)
        by_index = {v: k for k, v in self._speed_index.items()}
        cap = 0
        for i in range(n):
            m = int(lens[i])
            if sp is not None and sp[i] >= 0:
                m = tables.resampled_length(m, by_index[int(sp[i])])
            cap += (m + 7) // 8 * 8
        dst = np.zeros(max(cap, 8), dtype=np.int16)
        d_off = np.zeros(n + 1, dtype=np.int64)
        d_len = np.zeros(max(n, 1), dtype=np.int64)
        self._check(self._lib.fe_perturb(
            self._h, C.c_void_p(packed.ctypes.data), _ptr(off, C.c_int64), _ptr(lens, C.c_int64), n,
            None if sp is None else _ptr(sp, C.c_int32), None if gn is None else _ptr(gn, C.c_float),
            C.c_void_p(dst.ctypes.data), dst.size, _ptr(d_off, C.c_int64), _ptr(d_len, C.c_int64), None),
            "fe_perturb")
        return dst, d_off, d_len[:n]

    def perturb(self, pcm_list, speeds=None, gains=None):
        """Speed / volume perturbation only: int16 in -> list of int16 arrays."""
        if len(pcm_list) == 0:
            return []
        packed, off, lens = pack_pcm(pcm_list, np.int16)
        dst, d_off, d_len = self.perturb_packed(packed, off, lens, speeds, gains)
        return [dst[int(d_off[i]):int(d_off[i]) + int(d_len[i])].copy() for i in range(lens.size)]

    def postprocess(self, mats, mean=True, var=True, deltas=True, delta_mode=None):
        """CMVN and/or delta cube for a list of (L, D) float matrices (host in, host out):
        speechpy.processing.cmvn / speechpy.feature.extract_derivative_feature as batch calls."""
        if len(mats) == 0:
            return []
        D = int(mats[0].shape[1])
        mats = [np.ascontiguousarray(m, dtype=np.float32) for m in mats]
        if any(m.ndim != 2 or m.shape[1] != D for m in mats):
            raise ValueError("all matrices must be (L, %d)" % D)
        n = len(mats)
        nfr = np.asarray([m.shape[0] for m in mats], dtype=np.int32)
        sizes = (nfr.astype(np.int64) * D + 3) // 4 * 4
        off = np.zeros(n, dtype=np.int64)
        if n > 1:
            np.cumsum(sizes[:-1], out=off[1:])
        packed = np.zeros(max(int(sizes.sum()), 4), dtype=np.float32)
        for m, o in zip(mats, off):
            packed[o:o + m.size] = m.reshape(-1)
        W = 3 if deltas else 1
        out = np.empty(max(int(((nfr.astype(np.int64) * D * W + 3) // 4 * 4).sum()), 4), dtype=np.float32)
        out_off = np.zeros(n + 1, dtype=np.int64)
        mode = (_lib.FE_POST_MEAN if mean else 0) | (_lib.FE_POST_VAR if var else 0) | (_lib.FE_POST_DELTAS if deltas else 0)
        dm = self.config.delta_mode if delta_mode is None else delta_mode
        dm = {"speechpy_as_shipped": _lib.FE_DELTA_SPEECHPY, "time_regression": _lib.FE_DELTA_TIME_REGRESSION}[dm]
        self._check(self._lib.fe_postprocess(
            self._h, C.c_void_p(packed.ctypes.data), _ptr(off, C.c_int64), _ptr(nfr, C.c_int32), n, D, mode, dm,
            C.c_void_p(out.ctypes.data), out.size, _ptr(out_off, C.c_int64), None), "fe_postprocess")
        res = []
        for i in range(n):
            a = out[int(out_off[i]):int(out_off[i]) + int(nfr[i]) * D * W]
            res.append(a.reshape((int(nfr[i]), D, 3) if deltas else (int(nfr[i]), D)))
        return res

    # -- measurement hooks ---------------------------------------------------
    def sync(self):
        self._check(self._lib.fe_sync(self._h), "fe_sync")

    def set_profiling(self, on=True):
        self._lib.fe_set_profiling(self._h, int(bool(on)))

    def kernel_ms(self):
        ms = (C.c_float * 4)()
        self._check(self._lib.fe_get_kernel_ms(self._h, ms), "fe_get_kernel_ms")
        return {"resample": ms[0], "frames_to_statics": ms[1], "cmvn_delta_pack": ms[2], "device_pass": ms[3]}

    def measure_fp32_peak(self):
        """Best FP32 TFLOP/s any of the four FMA probes sustains on this GPU (roofline denominator)."""
        v = C.c_float()
        self._check(self._lib.fe_measure_fp32_peak(self._h, C.byref(v)), "fe_measure_fp32_peak")
        return float(v.value)

    def measure_fp32_peaks(self):
        """The four probes behind ``measure_fp32_peak`` (TFLOP/s each), keyed by instruction form."""
        v = (C.c_float * 4)()
        self._check(self._lib.fe_measure_fp32_peaks(self._h, v), "fe_measure_fp32_peaks")
        names = ("ffma_scalar_uniform_operands", "ffma_scalar_register_operands", "ffma2_packed_uniform_operands",
                 "ffma2_packed_register_operands")
        return {n: float(x) for n, x in zip(names, v)}

    def debug_counters(self):
        """K1 per-phase clock64 totals (needs FE_K1_DBG=8); profiling aid, see tools/k1_phases.py."""
        out = (C.c_uint64 * 16)()
        self._check(self._lib.fe_debug_counters(self._h, out), "fe_debug_counters")
        return [int(v) for v in out]

    def launch_count(self):
        return int(self._lib.fe_launch_count(self._h))

    def device_bytes(self):
        return int(self._lib.fe_device_bytes(self._h))
