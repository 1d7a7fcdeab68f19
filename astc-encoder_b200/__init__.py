"""astc-encoder_b200: B200-native ASTC block compressor behind the astcenc.h C ABI.

This module is the thin Python mirror of the C interface (same names and argument meaning as
astcenc.h: config_init -> context_alloc -> compress_image); all compression happens in
libastcenc_b200.so (hand-written sm_100a kernels). There is no CPU fallback: importing works
anywhere, but creating a context without a CUDA device raises.

PyTorch is used only as plumbing (device buffers, streams, torch.distributed); see
compress_device() for the device-resident entry point used by bench.py and the multi-GPU path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ASTCENC_B200_LIB") or os.path.join(_HERE, "libastcenc_b200.so")      # (the override is for A/B builds in tools/)

PRF_LDR_SRGB, PRF_LDR, PRF_HDR_RGB_LDR_A, PRF_HDR = 0, 1, 2, 3
PRE_FASTEST, PRE_FAST, PRE_MEDIUM, PRE_THOROUGH, PRE_VERYTHOROUGH, PRE_EXHAUSTIVE = 0.0, 10.0, 60.0, 98.0, 99.0, 100.0
FLG_MAP_NORMAL, FLG_USE_DECODE_UNORM8, FLG_USE_ALPHA_WEIGHT, FLG_USE_PERCEPTUAL = 1, 2, 4, 8
FLG_DECOMPRESS_ONLY, FLG_SELF_DECOMPRESS_ONLY, FLG_MAP_RGBM = 16, 32, 64
TYPE_U8, TYPE_F16, TYPE_F32 = 0, 1, 2
ERROR_NAMES = ["ASTCENC_SUCCESS", "ASTCENC_ERR_OUT_OF_MEM", "ASTCENC_ERR_BAD_CPU_FLOAT", "ASTCENC_ERR_BAD_PARAM", "ASTCENC_ERR_BAD_BLOCK_SIZE",
               "ASTCENC_ERR_BAD_PROFILE", "ASTCENC_ERR_BAD_QUALITY", "ASTCENC_ERR_BAD_SWIZZLE", "ASTCENC_ERR_BAD_FLAGS", "ASTCENC_ERR_BAD_CONTEXT",
               "ASTCENC_ERR_NOT_IMPLEMENTED", "ASTCENC_ERR_BAD_DECODE_MODE"]


ERR_OUT_OF_MEM, ERR_BAD_PARAM, ERR_BAD_BLOCK_SIZE, ERR_BAD_SWIZZLE, ERR_BAD_CONTEXT, ERR_NOT_IMPLEMENTED = 1, 3, 4, 7, 9, 10


class AstcencError(RuntimeError):
    def __init__(self, code, what):
        self.code = code
        RuntimeError.__init__(self, "%s: %s" % (what, ERROR_NAMES[code] if 0 <= code < len(ERROR_NAMES) else code))


class Config(C.Structure):
    """astcenc_config (astcenc.h:427-605)."""
    _fields_ = [
        ("profile", C.c_int), ("flags", C.c_uint), ("block_x", C.c_uint), ("block_y", C.c_uint), ("block_z", C.c_uint),
        ("cw_r_weight", C.c_float), ("cw_g_weight", C.c_float), ("cw_b_weight", C.c_float), ("cw_a_weight", C.c_float),
        ("a_scale_radius", C.c_uint), ("rgbm_m_scale", C.c_float),
        ("tune_partition_count_limit", C.c_uint), ("tune_2partition_index_limit", C.c_uint),
        ("tune_3partition_index_limit", C.c_uint), ("tune_4partition_index_limit", C.c_uint),
        ("tune_block_mode_limit", C.c_uint), ("tune_refinement_limit", C.c_uint), ("tune_candidate_limit", C.c_uint),
        ("tune_2partitioning_candidate_limit", C.c_uint), ("tune_3partitioning_candidate_limit", C.c_uint),
        ("tune_4partitioning_candidate_limit", C.c_uint),
        ("tune_db_limit", C.c_float), ("tune_mse_overshoot", C.c_float),
        ("tune_2partition_early_out_limit_factor", C.c_float), ("tune_3partition_early_out_limit_factor", C.c_float),
        ("tune_2plane_early_out_limit_correlation", C.c_float), ("tune_search_mode0_enable", C.c_float),
        ("progress_callback", C.c_void_p),
    ]


class Image(C.Structure):
    _fields_ = [("dim_x", C.c_uint), ("dim_y", C.c_uint), ("dim_z", C.c_uint), ("data_type", C.c_int), ("data", C.POINTER(C.c_void_p))]


class Swizzle(C.Structure):
    _fields_ = [("r", C.c_int), ("g", C.c_int), ("b", C.c_int), ("a", C.c_int)]


class ErrorMetrics(C.Structure):
    """struct astcenc_b200_error_metrics (include/astcenc.h)."""
    _fields_ = [(n, C.c_double) for n in ("psnr", "alpha_psnr", "rgb_psnr", "rgb_peak", "peak_psnr", "mpsnr", "log_rmse",
                                          "mean_angular_error", "worst_angular_error")] + [("sum_squared_error", C.c_double * 4)]


class CImageHeader(C.Structure):
    """struct astcenc_b200_cimage_header (include/astcenc.h): the fields of the 16-byte .astc header."""
    _fields_ = [(n, C.c_uint) for n in ("block_x", "block_y", "block_z", "dim_x", "dim_y", "dim_z")]


def build(verbose=False):
    """Compile libastcenc_b200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libastcenc_b200.so failed")
    return LIB_PATH


_lib = None


def lib():
    """The loaded C library. Fails loudly if it was not built - there is no Python/CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libastcenc_b200.so is missing: run `make -C %s` (or __graft_entry__.build())" % _HERE)
        l = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL)
        l.astcenc_config_init.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_float, C.c_uint, C.POINTER(Config)]
        l.astcenc_config_init.restype = C.c_int
        l.astcenc_context_alloc.argtypes = [C.POINTER(Config), C.c_uint, C.POINTER(C.c_void_p), C.c_void_p]
        l.astcenc_context_alloc.restype = C.c_int
        l.astcenc_compress_image.argtypes = [C.c_void_p, C.POINTER(Image), C.POINTER(Swizzle), C.c_void_p, C.c_size_t, C.c_uint]
        l.astcenc_compress_image.restype = C.c_int
        l.astcenc_compress_reset.argtypes = [C.c_void_p]
        l.astcenc_compress_reset.restype = C.c_int
        l.astcenc_decompress_image.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Image), C.POINTER(Swizzle), C.c_uint]
        l.astcenc_decompress_image.restype = C.c_int
        l.astcenc_compress_cancel.argtypes = [C.c_void_p]
        l.astcenc_compress_cancel.restype = C.c_int
        l.astcenc_context_free.argtypes = [C.c_void_p]
        l.astcenc_context_free.restype = None
        l.astcenc_get_error_string.argtypes = [C.c_int]
        l.astcenc_get_error_string.restype = C.c_char_p
        l.astcenc_b200_compress_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint, C.c_uint, C.POINTER(Swizzle), C.c_uint, C.c_uint, C.c_void_p, C.c_void_p]
        l.astcenc_b200_compress_device.restype = C.c_int
        l.astcenc_b200_launch_count.argtypes = [C.c_void_p]
        l.astcenc_b200_launch_count.restype = C.c_ulonglong
        l.astcenc_b200_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        l.astcenc_b200_last_timing.restype = C.c_int
        l.astcenc_b200_stage_timing.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_uint)]
        l.astcenc_b200_stage_timing.restype = C.c_int
        l.astcenc_b200_compute_error_metrics.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Image), C.POINTER(Image), C.c_int, C.c_int,
                                                         C.POINTER(ErrorMetrics)]
        l.astcenc_b200_compute_error_metrics.restype = C.c_int
        l.astcenc_b200_store_cimage.argtypes = [C.c_char_p, C.POINTER(CImageHeader), C.c_void_p, C.c_size_t]
        l.astcenc_b200_store_cimage.restype = C.c_int
        l.astcenc_b200_load_cimage.argtypes = [C.c_char_p, C.POINTER(CImageHeader), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        l.astcenc_b200_load_cimage.restype = C.c_int
        l.astcenc_b200_store_ktx_cimage.argtypes = [C.c_char_p, C.POINTER(CImageHeader), C.c_int, C.c_void_p, C.c_size_t]
        l.astcenc_b200_store_ktx_cimage.restype = C.c_int
        l.astcenc_b200_load_ktx_cimage.argtypes = [C.c_char_p, C.POINTER(CImageHeader), C.POINTER(C.c_int), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        l.astcenc_b200_load_ktx_cimage.restype = C.c_int
        l.astcenc_b200_comm_unique_id.argtypes = [C.c_void_p, C.c_size_t]
        l.astcenc_b200_comm_unique_id.restype = C.c_int
        l.astcenc_b200_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        l.astcenc_b200_comm_init.restype = C.c_int
        l.astcenc_b200_comm_free.argtypes = [C.c_void_p]
        l.astcenc_b200_comm_free.restype = C.c_int
        l.astcenc_b200_slab_rows.argtypes = [C.c_void_p, C.c_uint, C.c_int, C.c_int, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        l.astcenc_b200_slab_rows.restype = C.c_int
        l.astcenc_b200_compress_image_sharded.argtypes = [C.c_void_p, C.POINTER(Image), C.POINTER(Swizzle), C.c_void_p, C.c_size_t, C.c_int]
        l.astcenc_b200_compress_image_sharded.restype = C.c_int
        l.astcenc_b200_compress_batch.argtypes = [C.c_void_p, C.POINTER(C.POINTER(Image)), C.c_uint, C.POINTER(Swizzle), C.POINTER(C.c_void_p), C.c_size_t, C.c_int]
        l.astcenc_b200_compress_batch.restype = C.c_int
        l.astcenc_b200_comm_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        l.astcenc_b200_comm_last_timing.restype = C.c_int
        _lib = l
    return _lib


COMM_ID_BYTES = 128


def comm_unique_id():
    """astcenc_b200_comm_unique_id: the 128-byte NCCL rendezvous id (made on one rank, shipped to the others by the caller)."""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    err = lib().astcenc_b200_comm_unique_id(buf, COMM_ID_BYTES)
    if err:
        raise AstcencError(err, "astcenc_b200_comm_unique_id")
    return bytes(buf)


def config_init(profile, block_x, block_y, quality, flags=0, block_z=1, **overrides):
    """astcenc_config_init (astcenc.h:725-749); keyword overrides edit tune_* fields like the CLI's power-user switches."""
    cfg = Config()
    err = lib().astcenc_config_init(profile, block_x, block_y, block_z, quality, flags, C.byref(cfg))
    if err:
        raise AstcencError(err, "astcenc_config_init")
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


_DTYPES = {np.dtype(np.uint8): TYPE_U8, np.dtype(np.float16): TYPE_F16, np.dtype(np.float32): TYPE_F32}


class Context:
    """astcenc_context (astcenc_context_alloc / astcenc_compress_image / astcenc_context_free)."""

    def __init__(self, config, thread_count=1):
        self.config = config
        self.thread_count = thread_count
        self.handle = C.c_void_p()
        err = lib().astcenc_context_alloc(C.byref(config), thread_count, C.byref(self.handle), None)
        if err:
            self.handle = None
            raise AstcencError(err, "astcenc_context_alloc")

    def close(self):
        if self.handle:
            lib().astcenc_context_free(self.handle)
            self.handle = None

    __del__ = close

    def blocks(self, dim_x, dim_y):
        bx, by = self.config.block_x, self.config.block_y
        return (dim_x + bx - 1) // bx, (dim_y + by - 1) // by

    def compress_image(self, img, swizzle=(0, 1, 2, 3), out=None, thread_index=0):
        """Host-pointer path: img is a (H, W, 4) or (D, H, W, 4) numpy array of uint8 / float16 / float32."""
        img = np.ascontiguousarray(img)
        if img.ndim == 3:
            img = img[None]
        d, h, w = img.shape[:3]
        slices = (C.c_void_p * d)(*[img[z].ctypes.data for z in range(d)])
        image = Image(w, h, d, _DTYPES[img.dtype], slices)
        sw = Swizzle(*swizzle)
        nbx, nby = self.blocks(w, h)
        bz = max(1, self.config.block_z)      # 3D block sizes: blocks come out in (z, y, x) order
        if out is None:
            out = np.empty(nbx * nby * ((d + bz - 1) // bz) * 16, dtype=np.uint8)
        err = lib().astcenc_compress_image(self.handle, C.byref(image), C.byref(sw), out.ctypes.data, out.nbytes, thread_index)
        if err:
            raise AstcencError(err, "astcenc_compress_image")
        return out

    def decompress_image(self, blocks, dim_x, dim_y, dtype=np.uint8, swizzle=(0, 1, 2, 3), dim_z=1, thread_index=0):
        """astcenc_decompress_image: 16-byte blocks -> (H, W, 4) (or (D, H, W, 4)) array of uint8 / float16 / float32."""
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
        out = np.zeros((dim_z, dim_y, dim_x, 4), dtype=dtype)
        slices = (C.c_void_p * dim_z)(*[out[z].ctypes.data for z in range(dim_z)])
        image = Image(dim_x, dim_y, dim_z, _DTYPES[np.dtype(dtype)], slices)
        sw = Swizzle(*swizzle)
        err = lib().astcenc_decompress_image(self.handle, blocks.ctypes.data, blocks.nbytes, C.byref(image), C.byref(sw), thread_index)
        if err:
            raise AstcencError(err, "astcenc_decompress_image")
        return out[0] if dim_z == 1 else out

    def compress_device(self, d_pixels_ptr, data_type, dim_x, dim_y, d_out_ptr, block_row0=0, block_rows=None, swizzle=(0, 1, 2, 3), stream=0):
        """Device-resident path (astcenc_b200_compress_device): raw device pointers, enqueued on `stream`, no sync."""
        nbx, nby = self.blocks(dim_x, dim_y)
        if block_rows is None:
            block_rows = nby - block_row0
        sw = Swizzle(*swizzle)
        err = lib().astcenc_b200_compress_device(self.handle, C.c_void_p(d_pixels_ptr), data_type, dim_x, dim_y, C.byref(sw), block_row0, block_rows,
                                                 C.c_void_p(d_out_ptr), C.c_void_p(stream))
        if err:
            raise AstcencError(err, "astcenc_b200_compress_device")

    def launch_count(self):
        return int(lib().astcenc_b200_launch_count(self.handle))

    # ---- multi-GPU (one process per GPU; include/astcenc.h "Multi-GPU sharding") ----
    def comm_init(self, rank, world, unique_id=None):
        self.rank, self.world = rank, world
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(unique_id) if unique_id is not None else None
        err = lib().astcenc_b200_comm_init(self.handle, rank, world, buf, COMM_ID_BYTES if buf is not None else 0)
        if err:
            raise AstcencError(err, "astcenc_b200_comm_init")

    def comm_free(self):
        if self.handle:
            lib().astcenc_b200_comm_free(self.handle)

    def slab_rows(self, dim_y, rank, world):
        a = C.c_uint(); n = C.c_uint()
        err = lib().astcenc_b200_slab_rows(self.handle, dim_y, rank, world, C.byref(a), C.byref(n))
        if err:
            raise AstcencError(err, "astcenc_b200_slab_rows")
        return a.value, n.value

    def compress_image_sharded(self, img, out=None, root=0, swizzle=(0, 1, 2, 3)):
        """Slab mode (collective): every rank passes the same (H, W, 4) image; the root gets the whole payload in `out`."""
        img = np.ascontiguousarray(img)
        h, w = img.shape[:2]
        slices = (C.c_void_p * 1)(img.ctypes.data)
        image = Image(w, h, 1, _DTYPES[img.dtype], slices)
        sw = Swizzle(*swizzle)
        nbx, nby = self.blocks(w, h)
        is_root = getattr(self, "rank", 0) == root
        if out is None and is_root:
            out = np.empty(nbx * nby * 16, dtype=np.uint8)
        err = lib().astcenc_b200_compress_image_sharded(self.handle, C.byref(image), C.byref(sw), out.ctypes.data if out is not None else None,
                                                        out.nbytes if out is not None else 0, root)
        if err:
            raise AstcencError(err, "astcenc_b200_compress_image_sharded")
        return out

    def compress_batch(self, images, outs=None, root=0, swizzle=(0, 1, 2, 3)):
        """Batch mode (collective): images[i] (None for images of other ranks) -> outs[i] on the root."""
        n = len(images)
        keep = []
        ptrs = (C.POINTER(Image) * n)()
        for i, im in enumerate(images):
            if im is None:
                continue
            im = np.ascontiguousarray(im)
            slices = (C.c_void_p * 1)(im.ctypes.data)
            st = Image(im.shape[1], im.shape[0], 1, _DTYPES[im.dtype], slices)
            keep.append((im, slices, st))
            ptrs[i] = C.pointer(st)
        sw = Swizzle(*swizzle)
        optr = None
        each = 0
        if outs is not None:
            optr = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
            each = outs[0].nbytes
        err = lib().astcenc_b200_compress_batch(self.handle, ptrs, n, C.byref(sw), optr, each, root)
        if err:
            raise AstcencError(err, "astcenc_b200_compress_batch")
        return outs

    def comm_last_timing(self):
        a = C.c_float(); b = C.c_float()
        lib().astcenc_b200_comm_last_timing(self.handle, C.byref(a), C.byref(b))
        return a.value, b.value

    def stage_timing(self, enable, fetch=False):
        """Enable / disable per-launch events; with fetch=True also return ({kernel: ms}, {kernel: launches}) of the last call."""
        names = ("setup", "refine", "prepare", "emit")
        if fetch:
            ms = (C.c_float * 4)(); n = (C.c_uint * 4)()
            lib().astcenc_b200_stage_timing(self.handle, 1 if enable else 0, ms, n)
            return {k: float(ms[i]) for i, k in enumerate(names)}, {k: int(n[i]) for i, k in enumerate(names)}
        lib().astcenc_b200_stage_timing(self.handle, 1 if enable else 0, None, None)
        return None

    def compute_error_metrics(self, img1, img2, hdr=False, normal=False, input_components=4, fstop_lo=-10, fstop_hi=10):
        """astcenc_b200_compute_error_metrics: the CLI's compute_error_metrics (astcenccli_error_metrics.cpp:109-413) on the
        device. img1 = original, img2 = decoded, numpy (H, W, 4) uint8 / float16 / float32. Returns a dict."""
        imgs = []
        for im in (img1, img2):
            im = np.ascontiguousarray(im)
            slices = (C.c_void_p * 1)(im.ctypes.data)
            imgs.append((im, slices, Image(im.shape[1], im.shape[0], 1, _DTYPES[im.dtype], slices)))
        out = ErrorMetrics()
        err = lib().astcenc_b200_compute_error_metrics(self.handle, int(hdr), int(normal), input_components, C.byref(imgs[0][2]), C.byref(imgs[1][2]),
                                                       fstop_lo, fstop_hi, C.byref(out))
        if err:
            raise AstcencError(err, "astcenc_b200_compute_error_metrics")
        d = {n: getattr(out, n) for n, _ in ErrorMetrics._fields_[:9]}
        d["sum_squared_error"] = list(out.sum_squared_error)
        return d

    def last_timing(self):
        ms = C.c_float(); a = C.c_size_t(); b = C.c_size_t()
        lib().astcenc_b200_last_timing(self.handle, C.byref(ms), C.byref(a), C.byref(b))
        return ms.value, a.value, b.value


def store_cimage(filename, blocks, dim_x, dim_y, block_x, block_y, dim_z=1, block_z=1):
    """Write a .astc file (store_cimage, astcenccli_image_load_store.cpp:2691-2730): 16-byte header + the blocks."""
    data = np.ascontiguousarray(np.frombuffer(bytes(blocks), dtype=np.uint8) if not isinstance(blocks, np.ndarray) else blocks.view(np.uint8).reshape(-1))
    hdr = CImageHeader(block_x, block_y, block_z, dim_x, dim_y, dim_z)
    err = lib().astcenc_b200_store_cimage(os.fsencode(filename), C.byref(hdr), data.ctypes.data, data.nbytes)
    if err:
        raise AstcencError(err, "astcenc_b200_store_cimage")


def load_cimage(filename):
    """Read a .astc file (load_cimage, astcenccli_image_load_store.cpp:2599-2688). Returns (blocks uint8 array, header dict);
    corrupt files raise AstcencError like the reference tool refuses them."""
    hdr = CImageHeader()
    n = C.c_size_t()
    err = lib().astcenc_b200_load_cimage(os.fsencode(filename), C.byref(hdr), None, 0, C.byref(n))
    if err:
        raise AstcencError(err, "astcenc_b200_load_cimage")
    data = np.empty(n.value, dtype=np.uint8)
    err = lib().astcenc_b200_load_cimage(os.fsencode(filename), C.byref(hdr), data.ctypes.data, data.nbytes, C.byref(n))
    if err:
        raise AstcencError(err, "astcenc_b200_load_cimage")
    return data, {k: getattr(hdr, k) for k, _ in CImageHeader._fields_}


def store_ktx_cimage(filename, blocks, dim_x, dim_y, block_x, block_y, is_srgb=False, dim_z=1, block_z=1):
    """Write the blocks as a KTX 1 file (store_ktx_compressed_image, astcenccli_image_load_store.cpp:1396-1440)."""
    data = np.ascontiguousarray(np.frombuffer(bytes(blocks), dtype=np.uint8) if not isinstance(blocks, np.ndarray) else blocks.view(np.uint8).reshape(-1))
    hdr = CImageHeader(block_x, block_y, block_z, dim_x, dim_y, dim_z)
    err = lib().astcenc_b200_store_ktx_cimage(os.fsencode(filename), C.byref(hdr), int(is_srgb), data.ctypes.data, data.nbytes)
    if err:
        raise AstcencError(err, "astcenc_b200_store_ktx_cimage")


def load_ktx_cimage(filename):
    """Read an ASTC payload from a KTX 1 file (load_ktx_compressed_image, :1294-1385). Returns (blocks, header dict, is_srgb)."""
    hdr = CImageHeader()
    n = C.c_size_t()
    srgb = C.c_int()
    err = lib().astcenc_b200_load_ktx_cimage(os.fsencode(filename), C.byref(hdr), C.byref(srgb), None, 0, C.byref(n))
    if err:
        raise AstcencError(err, "astcenc_b200_load_ktx_cimage")
    data = np.empty(n.value, dtype=np.uint8)
    err = lib().astcenc_b200_load_ktx_cimage(os.fsencode(filename), C.byref(hdr), C.byref(srgb), data.ctypes.data, data.nbytes, C.byref(n))
    if err:
        raise AstcencError(err, "astcenc_b200_load_ktx_cimage")
    return data, {k: getattr(hdr, k) for k, _ in CImageHeader._fields_}, bool(srgb.value)


def slab_rows(blocks_y, rank, world):
    """Block-row slab [r0, r1) of `rank` among `world` ranks."""
    r0 = (blocks_y * rank) // world
    r1 = (blocks_y * (rank + 1)) // world
    return r0, r1
