"""Minimal UVH5 reader for the driver (`run_hydra_pspec_b200.py`).

The reference reads its input through pyuvdata (`run-hydra-pspec.py:305-322`); pyuvdata and h5py are
not part of this image, so the driver falls back to this pure-python reader when they cannot be
imported.  It understands exactly the HDF5 subset that h5py's default (`libver='earliest'`) writer
produces for UVH5 files: version-0 superblock, version-1 object headers, symbol-table groups
(v1 B-tree + local heap), contiguous / compact / chunked (v1 B-tree) dataset layouts, the deflate,
shuffle and LZF filters, and fixed-point, floating-point, fixed-length string and compound
(`r`, `i`) datatypes.  Anything else raises ``NotImplementedError``.

Only what the hot path's caller needs is exposed: the header arrays, `visdata`, `flags`, `nsamples`,
and the few `pyuvdata.UVData` operations `run-hydra-pspec.py` uses (`select` by antenna pairs and
frequencies, `conjugate_bls`, `get_antpairs`, `get_data`, `get_flags`, pseudo-Stokes I).
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _lzf_decompress(src, out_len):
    """LZF (liblzf) decoder, as used by h5py's filter 32000."""
    out = bytearray(out_len)
    ip, op, n = 0, 0, len(src)
    while ip < n:
        ctrl = src[ip]
        ip += 1
        if ctrl < 32:  # literal run
            ln = ctrl + 1
            out[op:op + ln] = src[ip:ip + ln]
            ip += ln
            op += ln
        else:  # back reference
            ln = ctrl >> 5
            ref = op - ((ctrl & 0x1F) << 8) - 1
            if ln == 7:
                ln += src[ip]
                ip += 1
            ref -= src[ip]
            ip += 1
            ln += 2
            for _ in range(ln):  # may overlap
                out[op] = out[ref]
                op += 1
                ref += 1
    if op != out_len:
        raise ValueError("LZF: decoded length mismatch")
    return bytes(out)


class _H5File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        if self.b[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file")
        if self.b[8] not in (0, 1):
            raise NotImplementedError("only version 0/1 HDF5 superblocks are supported")
        if self.b[13] != 8 or self.b[14] != 8:
            raise NotImplementedError("only 8-byte offsets/lengths are supported")
        off = 24 if self.b[8] == 0 else 28
        self.base = struct.unpack_from("<Q", self.b, off)[0]
        root_entry = off + 32
        self.root = struct.unpack_from("<Q", self.b, root_entry + 8)[0]

    # ---- object headers (version 1)
    def messages(self, addr):
        b = self.b
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise NotImplementedError("only version-1 object headers are supported")
        out = []
        blocks = [(addr + 16, hsize)]
        while blocks:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                body = pos + 8
                if mtype == 0x10:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", b, body)
                    blocks.append((caddr, clen))
                out.append((mtype, body, msize))
                pos = body + msize
        return out

    # ---- groups (symbol table)
    def group_entries(self, addr):
        for mtype, body, _ in self.messages(addr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", self.b, body)
                return self._walk_group_btree(btree, self._heap_data(heap))
        raise NotImplementedError("group without a symbol-table message (new-style groups are not supported)")

    def _heap_data(self, addr):
        if self.b[addr:addr + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        return struct.unpack_from("<Q", self.b, addr + 24)[0]

    def _walk_group_btree(self, addr, heap_data):
        b = self.b
        if b[addr:addr + 4] == b"SNOD":
            nsym = struct.unpack_from("<H", b, addr + 6)[0]
            ents = {}
            for i in range(nsym):
                e = addr + 8 + 40 * i
                noff, ohdr = struct.unpack_from("<QQ", b, e)
                s = heap_data + noff
                name = b[s:b.index(b"\x00", s)].decode()
                ents[name] = ohdr
            return ents
        if b[addr:addr + 4] != b"TREE":
            raise ValueError("bad B-tree signature")
        _ntype, _level, nent = struct.unpack_from("<BBH", b, addr + 4)
        ents = {}
        pos = addr + 24
        for i in range(nent):
            child = struct.unpack_from("<Q", b, pos + 8 + 16 * i)[0]
            ents.update(self._walk_group_btree(child, heap_data))
        return ents

    # ---- datasets
    def _parse_dtype(self, body):
        b = self.b
        cv = b[body]
        cls, ver = cv & 0x0F, cv >> 4
        bits0 = b[body + 1]
        size = struct.unpack_from("<I", b, body + 4)[0]
        if cls == 0:  # fixed point
            signed = bool(bits0 & 0x08)
            return np.dtype(("<" if not (bits0 & 1) else ">") + ("i" if signed else "u") + str(size)), 8 + 4
        if cls == 1:  # float
            return np.dtype(("<" if not (bits0 & 1) else ">") + "f" + str(size)), 8 + 12
        if cls == 3:  # string
            return np.dtype("S" + str(size)), 8
        if cls == 6:  # compound
            nmemb = struct.unpack_from("<H", b, body + 1)[0]
            pos = body + 8
            names, formats, offsets = [], [], []
            for _ in range(nmemb):
                end = b.index(b"\x00", pos)
                name = b[pos:end].decode()
                if ver < 3:
                    pos += ((end - pos) // 8 + 1) * 8
                    moff = struct.unpack_from("<I", b, pos)[0]
                    pos += 4
                    if ver == 1:
                        pos += 1 + 3 + 4 + 4 + 16
                else:
                    pos = end + 1
                    nb = max(1, (size.bit_length() + 7) // 8)
                    moff = int.from_bytes(b[pos:pos + nb], "little")
                    pos += nb
                mdt, mlen = self._parse_dtype(pos)
                pos += mlen
                names.append(name)
                formats.append(mdt)
                offsets.append(moff)
            return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size}), pos - body
        if cls == 8:  # enum (h5py stores bool as enum of int8)
            base, blen = self._parse_dtype(body + 8)
            return base, 8 + blen  # member list not needed
        raise NotImplementedError(f"HDF5 datatype class {cls}")

    def read_dataset(self, addr):
        b = self.b
        shape, dtype, layout, filters = (), None, None, []
        for mtype, body, msize in self.messages(addr):
            if mtype == 0x01:  # dataspace
                ver, rank, flags = struct.unpack_from("<BBB", b, body)
                pos = body + (8 if ver == 1 else 4)
                shape = struct.unpack_from("<" + "Q" * rank, b, pos) if rank else ()
            elif mtype == 0x03:
                dtype, _ = self._parse_dtype(body)
            elif mtype == 0x08:
                ver, cls = struct.unpack_from("<BB", b, body)
                if ver != 3:
                    raise NotImplementedError("only version-3 data layout messages are supported")
                if cls == 1:
                    layout = ("contiguous",) + struct.unpack_from("<QQ", b, body + 2)
                elif cls == 2:
                    rank = b[body + 2]
                    bt = struct.unpack_from("<Q", b, body + 3)[0]
                    dims = struct.unpack_from("<" + "I" * rank, b, body + 11)
                    layout = ("chunked", bt, dims)
                elif cls == 0:
                    sz = struct.unpack_from("<H", b, body + 2)[0]
                    layout = ("compact", body + 4, sz)
            elif mtype == 0x0B:  # filter pipeline
                ver, nf = struct.unpack_from("<BB", b, body)
                pos = body + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid, nlen, _fl, ncd = struct.unpack_from("<HHHH", b, pos)
                    pos += 8
                    if ver == 1 or fid >= 256:
                        pos += (nlen + 7) // 8 * 8 if ver == 1 else nlen
                    cd = struct.unpack_from("<" + "I" * ncd, b, pos)
                    pos += 4 * ncd
                    if ver == 1 and ncd % 2:
                        pos += 4
                    filters.append((fid, cd))
        if dtype is None or layout is None:
            raise ValueError("dataset without datatype/layout")
        count = int(np.prod(shape)) if shape else 1
        if layout[0] == "contiguous":
            a, sz = layout[1], layout[2]
            if a == UNDEF:
                return np.zeros(shape, dtype)
            return np.frombuffer(b, dtype=dtype, count=count, offset=a + self.base).reshape(shape).copy()
        if layout[0] == "compact":
            return np.frombuffer(b, dtype=dtype, count=count, offset=layout[1]).reshape(shape).copy()
        _, bt, cdims = layout
        cshape = cdims[:-1]
        out = np.zeros(shape, dtype)
        if bt != UNDEF:
            for offs, caddr, csize, fmask in self._walk_chunk_btree(bt, len(cdims)):
                raw = b[caddr:caddr + csize]
                # bit i of the chunk's filter mask set = filter i of the pipeline was skipped for this chunk
                # (optional filters such as LZF leave incompressible chunks unfiltered)
                for fidx in range(len(filters) - 1, -1, -1):
                    if (fmask >> fidx) & 1:
                        continue
                    fid, cd = filters[fidx]
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 32000:
                        raw = _lzf_decompress(raw, cd[2] if len(cd) > 2 else int(np.prod(cshape)) * dtype.itemsize)
                    elif fid == 2:
                        es = cd[0]
                        arr = np.frombuffer(raw, np.uint8).reshape(es, -1)
                        raw = arr.T.tobytes()
                    else:
                        raise NotImplementedError(f"HDF5 filter {fid}")
                chunk = np.frombuffer(raw, dtype=dtype, count=int(np.prod(cshape))).reshape(cshape)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, shape))
                out[sl] = chunk[tuple(slice(0, x.stop - x.start) for x in sl)]
        return out

    def _walk_chunk_btree(self, addr, nd):
        b = self.b
        if b[addr:addr + 4] != b"TREE":
            raise ValueError("bad chunk B-tree signature")
        ntype, level, nent = struct.unpack_from("<BBH", b, addr + 4)
        if ntype != 1:
            raise ValueError("expected a raw-data chunk B-tree")
        pos = addr + 24
        keysize = 8 + 8 * nd
        for i in range(nent):
            k = pos + i * (keysize + 8)
            csize, fmask = struct.unpack_from("<II", b, k)
            offs = struct.unpack_from("<" + "Q" * nd, b, k + 8)[:-1]
            child = struct.unpack_from("<Q", b, k + keysize)[0]
            if level == 0:
                yield offs, child + self.base, csize, fmask
            else:
                yield from self._walk_chunk_btree(child, nd)


def _to_python(a):
    if a.dtype.kind == "S":
        return a.tobytes().rstrip(b"\x00").decode() if a.shape == () else np.char.decode(a)
    if a.shape == ():
        return a.item()
    return a


class UVH5Data:
    """The slice of `pyuvdata.UVData` that `run-hydra-pspec.py` touches."""

    def __init__(self, path):
        h5 = _H5File(path)
        root = h5.group_entries(h5.root)
        hdr = h5.group_entries(root["Header"])
        dat = h5.group_entries(root["Data"])
        self.header = {}
        for k, a in hdr.items():
            try:
                self.header[k] = _to_python(h5.read_dataset(a))
            except (NotImplementedError, ValueError):
                continue  # sub-groups / exotic entries (extra_keywords) are not needed
        vis = h5.read_dataset(dat["visdata"])
        if vis.dtype.names:
            vis = vis["r"].astype(np.float64) + 1j * vis["i"].astype(np.float64)
        flags = h5.read_dataset(dat["flags"]).astype(bool)
        nsamples = h5.read_dataset(dat["nsamples"]).astype(np.float64)
        if vis.ndim == 4:  # (Nblts, Nspws = 1, Nfreqs, Npols): pre-"future shapes" files
            vis, flags, nsamples = vis[:, 0], flags[:, 0], nsamples[:, 0]
        self.data_array = np.ascontiguousarray(vis, dtype=np.complex128)
        self.flag_array = flags
        self.nsample_array = nsamples
        self.ant_1_array = np.asarray(self.header["ant_1_array"]).astype(int)
        self.ant_2_array = np.asarray(self.header["ant_2_array"]).astype(int)
        self.time_array = np.asarray(self.header["time_array"], dtype=float)
        fa = np.asarray(self.header["freq_array"], dtype=float)
        self.freq_array = fa.reshape(-1)
        self.polarization_array = np.asarray(self.header["polarization_array"]).astype(int).reshape(-1)
        self.Nfreqs = self.freq_array.size

    # -- pyuvdata.UVData.select(ant_str=..., frequencies=...)
    def select(self, ant_str="cross", frequencies=None):
        keep = np.ones(self.ant_1_array.size, dtype=bool)
        if ant_str in ("cross", None):
            keep &= self.ant_1_array != self.ant_2_array
        elif ant_str == "auto":
            keep &= self.ant_1_array == self.ant_2_array
        elif ant_str != "all":
            pairs = set()
            for tok in ant_str.split(","):
                a, b = tok.split("_")
                pairs.add((int(a), int(b)))
                pairs.add((int(b), int(a)))
            keep &= np.array([(a, b) in pairs for a, b in zip(self.ant_1_array, self.ant_2_array)])
        self.data_array, self.flag_array, self.nsample_array = (x[keep] for x in (self.data_array, self.flag_array,
                                                                                  self.nsample_array))
        self.ant_1_array, self.ant_2_array, self.time_array = (x[keep] for x in (self.ant_1_array, self.ant_2_array,
                                                                                 self.time_array))
        if frequencies is not None:
            fk = np.array([np.argmin(np.abs(self.freq_array - f)) for f in np.atleast_1d(frequencies)])
            fk = np.unique(fk)
            self.data_array, self.flag_array, self.nsample_array = (x[:, fk] for x in (self.data_array, self.flag_array,
                                                                                       self.nsample_array))
            self.freq_array = self.freq_array[fk]
            self.Nfreqs = fk.size

    # -- pyuvdata.UVData.conjugate_bls() (default convention "ant1<ant2")
    def conjugate_bls(self):
        sw = self.ant_1_array > self.ant_2_array
        self.data_array[sw] = np.conj(self.data_array[sw])
        # conjugating a baseline swaps its cross-hand polarisations (xy <-> yx, rl <-> lr), as pyuvdata does
        pols = list(self.polarization_array)
        for pa, pb in ((-7, -8), (-3, -4)):
            if pa in pols and pb in pols:
                ia, ib = pols.index(pa), pols.index(pb)
                for arr in (self.data_array, self.flag_array, self.nsample_array):
                    tmp = arr[sw][..., ia].copy()
                    arr[sw, ..., ia] = arr[sw][..., ib]
                    arr[sw, ..., ib] = tmp
        a1 = self.ant_1_array.copy()
        self.ant_1_array[sw] = self.ant_2_array[sw]
        self.ant_2_array[sw] = a1[sw]

    def get_antpairs(self):
        seen, out = set(), []
        for a, b in zip(self.ant_1_array, self.ant_2_array):
            if (a, b) not in seen:
                seen.add((a, b))
                out.append((int(a), int(b)))
        return out

    def _pol_index(self, pol):
        num = {"xx": -5, "yy": -6, "xy": -7, "yx": -8, "pI": 1}[pol]
        idx = np.nonzero(self.polarization_array == num)[0]
        if idx.size == 0:
            raise KeyError(pol)
        return int(idx[0])

    def _rows(self, antpair):
        rows = np.nonzero((self.ant_1_array == antpair[0]) & (self.ant_2_array == antpair[1]))[0]
        return rows[np.argsort(self.time_array[rows], kind="stable")]

    def get_data(self, key):
        """(Ntimes, Nfreqs) visibilities of baseline (ant1, ant2, pol); always a copy."""
        return self.data_array[self._rows(key[:2])][:, :, self._pol_index(key[2])].copy()

    def get_flags(self, key):
        return self.flag_array[self._rows(key[:2])][:, :, self._pol_index(key[2])].copy()

    def get_nsamples(self, key):
        return self.nsample_array[self._rows(key[:2])][:, :, self._pol_index(key[2])].copy()

    # -- hydra_pspec.utils.form_pseudo_stokes_vis (utils.py:104-135)
    def form_pseudo_stokes_vis(self, convention=1.0):
        if 1 in self.polarization_array:
            return
        ix, iy = self._pol_index("xx"), self._pol_index("yy")
        self.data_array[..., ix] += self.data_array[..., iy]
        self.data_array *= convention
        self.data_array = self.data_array[..., ix:ix + 1]
        self.flag_array = self.flag_array[..., ix:ix + 1]
        self.nsample_array = self.nsample_array[..., ix:ix + 1]
        self.polarization_array = np.array([-5])


def read_uvh5(path):
    """One file, or a list of files with the same frequency / polarisation axes concatenated along the
    baseline-time axis (what ``UVData.read([...])`` does for files that split an observation in time)."""
    if isinstance(path, (list, tuple)):
        parts = [UVH5Data(p) for p in path]
        uv = parts[0]
        for o in parts[1:]:
            if o.freq_array.shape != uv.freq_array.shape or not np.allclose(o.freq_array, uv.freq_array) \
                    or list(o.polarization_array) != list(uv.polarization_array):
                raise ValueError("uvh5 files to concatenate must share their frequency and polarisation axes")
        uv.data_array = np.concatenate([o.data_array for o in parts])
        uv.flag_array = np.concatenate([o.flag_array for o in parts])
        uv.nsample_array = np.concatenate([o.nsample_array for o in parts])
        uv.ant_1_array = np.concatenate([o.ant_1_array for o in parts])
        uv.ant_2_array = np.concatenate([o.ant_2_array for o in parts])
        uv.time_array = np.concatenate([o.time_array for o in parts])
        return uv
    return UVH5Data(path)
