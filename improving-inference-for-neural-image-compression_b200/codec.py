"""Entropy coding of the quantised latents and the bitstream container -- the `string` / `side_string` side of the
reference's mbt2018.py (:84-85 compress, :211-221 pack, :248-295 decompress), where the reference calls
tensorflow-compression's range-coding ops and `tfc.PackedTensors`.

Device work (table build, symbol binning, h_s for the decoder) and the host range coder live in libsga_b200.so
(csrc/codec.cu); this module only moves arrays and packs bytes.  The container is this repo's own (".tfci" by name only):

    magic "SGAC" | u8 version | u8 mode | u16 C | u32 N | u32 H | u32 W | N x (u32 len_z, u32 len_y) | streams (z_0 y_0 z_1 y_1 ...)

mode 0: y_hat = round(y) (the result of sga.py:240-241);  mode 1: y_hat = round(y - mu) + mu (mbt2018.py baseline)."""
import ctypes
import struct

import numpy as np
import torch

from . import _lib

MAGIC = b"SGAC"
VERSION = 1
MODE_SGA = 0
MODE_MEAN_CENTRED = 1


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _np_ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


class HostCodec:
    """The range coder over given tables (no GPU needed): used by the tests and by LatentCodec."""

    def __init__(self, handle, lib):
        self.lib, self.h = lib, handle

    @classmethod
    def from_tables(cls, tables_y, tables_z):
        """tables_*: list of (min_symbol, frequencies) with frequencies = regular symbols then the escape symbol, summing to
        65536."""
        lib = _lib.load()

        def pack(tabs):
            ln = np.asarray([len(f) - 1 for _, f in tabs], dtype=np.int32)
            mn = np.asarray([m for m, _ in tabs], dtype=np.int32)
            fr = np.concatenate([np.asarray(f, dtype=np.uint32) for _, f in tabs]) if tabs else np.zeros(0, np.uint32)
            return ln, mn, fr
        ly, my, fy = pack(tables_y)
        lz, mz, fz = pack(tables_z)
        h = ctypes.c_void_p()
        _lib.check(lib.sga_codec_create_host(len(tables_y), _np_ptr(ly), _np_ptr(my), _np_ptr(fy), len(tables_z), _np_ptr(lz),
                                             _np_ptr(mz), _np_ptr(fz), ctypes.byref(h)))
        return cls(h, lib)

    def close(self):
        if self.h is not None and self.h.value:
            self.lib.sga_codec_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode(self, which, tab, sym):
        tab = np.ascontiguousarray(tab, dtype=np.int32)
        sym = np.ascontiguousarray(sym, dtype=np.int32)
        assert tab.shape == sym.shape
        n = int(tab.size)
        cap = max(64, n * 2 + 64)
        while True:
            out = np.empty(cap, dtype=np.uint8)
            nb = ctypes.c_uint64(0)
            rc = self.lib.sga_rc_encode(self.h, int(which), n, _np_ptr(tab), _np_ptr(sym), _np_ptr(out), cap, ctypes.byref(nb))
            if rc != 0 and nb.value > cap:
                cap = int(nb.value)
                continue
            _lib.check(rc)
            return out[:nb.value].tobytes()

    def decode(self, which, tab, data):
        tab = np.ascontiguousarray(tab, dtype=np.int32)
        buf = np.frombuffer(data, dtype=np.uint8)
        sym = np.empty(tab.shape, dtype=np.int32)
        _lib.check(self.lib.sga_rc_decode(self.h, int(which), int(tab.size), _np_ptr(tab), _np_ptr(buf) if buf.size else None,
                                          int(buf.size), _np_ptr(sym)))
        return sym


class LatentCodec(HostCodec):
    """Tables of one bound engine (Gaussian-conditional tables + the factorised prior of its weights)."""

    def __init__(self, engine):
        self.eng = engine
        lib = engine.lib
        h = ctypes.c_void_p()
        _lib.check(lib.sga_codec_create(engine.ctx, ctypes.byref(h), _stream()))
        super().__init__(h, lib)
        c = np.empty(engine.c, dtype=np.int32)
        _lib.check(lib.sga_codec_z_centers(self.h, _np_ptr(c), engine.c))
        self.z_centers = c

    def _bin(self, y, z, mode):
        e = self.eng
        dev = e.device
        tab_y = torch.empty(e.y_shape, dtype=torch.int32, device=dev)
        base_y = torch.empty(e.y_shape, dtype=torch.float32, device=dev)
        sym_y = torch.empty(e.y_shape, dtype=torch.int32, device=dev) if y is not None else None
        yhat = torch.empty(e.y_shape, dtype=torch.float32, device=dev) if y is not None else None
        tab_z = torch.empty(e.z_shape, dtype=torch.int32, device=dev) if y is not None else None
        sym_z = torch.empty(e.z_shape, dtype=torch.int32, device=dev) if y is not None else None
        _lib.check(self.lib.sga_codec_bin(e.ctx, self.h, _ptr(y), _ptr(z), int(mode), _ptr(tab_y), _ptr(sym_y), _ptr(base_y),
                                          _ptr(yhat), _ptr(tab_z), _ptr(sym_z), _stream()))
        return tab_y, sym_y, base_y, yhat, tab_z, sym_z

    def quantise(self, y, z, mode):
        """(y_hat, z_hat) as the bitstream will reproduce them."""
        _, _, _, yhat, _, _ = self._bin(y, z, mode)
        return yhat, torch.round(z)

    def compress(self, y, z, mode=MODE_SGA):
        """y, z: device tensors of the engine's latent shapes.  Returns (list of (z_bytes, y_bytes) per image, y_hat, z_hat)."""
        tab_y, sym_y, _, yhat, tab_z, sym_z = self._bin(y, z, mode)
        n = self.eng.n
        ty, sy = tab_y.cpu().numpy().reshape(n, -1), sym_y.cpu().numpy().reshape(n, -1)
        tz, sz = tab_z.cpu().numpy().reshape(n, -1), sym_z.cpu().numpy().reshape(n, -1)
        streams = [(self.encode(1, tz[i], sz[i]), self.encode(0, ty[i], sy[i])) for i in range(n)]
        return streams, yhat, torch.round(z)

    def decompress(self, streams, mode=MODE_SGA):
        """Inverse of compress: (y_hat, z_hat) device tensors, bit-identical to what the encoder quantised."""
        e = self.eng
        n, C = e.n, e.c
        assert len(streams) == n
        per_z = int(np.prod(e.z_shape[1:]))
        tab_z = np.tile(np.arange(C, dtype=np.int32), per_z // C)
        zs = np.stack([self.decode(1, tab_z, streams[i][0]) for i in range(n)])
        z_hat = (zs + self.z_centers[tab_z][None, :]).astype(np.float32).reshape(e.z_shape)
        z_dev = torch.from_numpy(z_hat).to(e.device)
        tab_y, _, base_y, _, _, _ = self._bin(None, z_dev, mode)
        ty, by = tab_y.cpu().numpy().reshape(n, -1), base_y.cpu().numpy().reshape(n, -1)
        ys = np.stack([self.decode(0, ty[i], streams[i][1]) for i in range(n)])
        y_hat = (by + ys.astype(np.float32)).astype(np.float32).reshape(e.y_shape)
        return torch.from_numpy(y_hat).to(e.device), z_dev


def pack(streams, mode, C, H, W):
    n = len(streams)
    head = MAGIC + struct.pack("<BBHIII", VERSION, mode, C, n, H, W)
    lens = b"".join(struct.pack("<II", len(z), len(y)) for z, y in streams)
    return head + lens + b"".join(z + y for z, y in streams)


def unpack(data):
    """Returns (streams, mode, C, H, W)."""
    if data[:4] != MAGIC:
        raise ValueError("not an SGAC bitstream (magic %r)" % data[:4])
    version, mode, C, n, H, W = struct.unpack_from("<BBHIII", data, 4)
    if version != VERSION:
        raise ValueError("unsupported bitstream version %d" % version)
    off = 4 + struct.calcsize("<BBHIII")
    lens = [struct.unpack_from("<II", data, off + 8 * i) for i in range(n)]
    off += 8 * n
    streams = []
    for lz, ly in lens:
        if off + lz + ly > len(data):
            raise ValueError("truncated bitstream")
        streams.append((data[off:off + lz], data[off + lz:off + lz + ly]))
        off += lz + ly
    return streams, mode, C, H, W
