"""Output writer with the reference's ALPSCore HDF5 layout (prog/data_save.hxx:33-151, prog/data_save.hpp:124-156).

libhdf5 / h5py do not exist in this image, so the container format is emitted directly: HDF5 file-format
specification 1.x with the oldest (most widely readable) structures -- version-0 superblock, version-1 object
headers, symbol-table groups (v1 B-tree node + local heap + one symbol-table node per group) and contiguous
little-endian datasets.  That is the same subset libhdf5 writes by default ("earliest" library bounds), so the
reference's consumers (h5py in scripts/parse/parse_thermod.py:48-49) read the files unchanged.

`H5Reader` parses the same subset; tests pin it on a file written by the real library (a MATLAB-7.3 fixture
shipped with scipy) and then use it to check the writer.

Layout written by `save_all_data`:
    /parameters/<name>                      scalars / strings (alps::params dump)
    /mc_data/{energies,d2energies,c_energies}   1-D float64 raw series (data_save.hxx:93-103)
    /mc_data/{ipr_history,spectrum_history,focc_history}  2-D [index][measurement] when given (:105-137)
    /binning/<obs>                          nbins x 5 rows [n, mean, variance, stderr, tau_int] (data_save.hpp:139-152)
    /stats/<obs>                            4-vector [n, mean, variance, stderr] at the plateau bin (:124-135)
with <obs> in {energy, d2energy, c_energy, cv}.
"""
import struct

import numpy as np

from . import stats

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K = 4         # libhdf5 defaults: a symbol-table node holds 2 * LEAF_K entries,
INTERNAL_K = 16    # a B-tree node 2 * INTERNAL_K children


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


class _Group:
    def __init__(self):
        self.children = {}   # name -> _Group | np.ndarray | bytes


class H5Writer:
    """Collects a tree of groups / datasets and serialises it in one pass.

        w = H5Writer(); w["/stats/energy"] = np.array([...]); w["/parameters/beta"] = 10.0; w.save("out.h5")
    """

    def __init__(self):
        self.root = _Group()

    def __setitem__(self, path, value):
        parts = [p for p in path.split("/") if p]
        if not parts:
            raise ValueError("empty dataset path")
        g = self.root
        for p in parts[:-1]:
            nxt = g.children.setdefault(p, _Group())
            if not isinstance(nxt, _Group):
                raise ValueError("%s is a dataset, not a group" % p)
            g = nxt
        if isinstance(value, str):
            value = value.encode()
        if isinstance(value, bytes):
            g.children[parts[-1]] = value.rstrip(b"\0") + b"\0"   # fixed-length, null-terminated: the terminator is part of the size
            return
        a = np.asarray(value)
        if a.dtype == np.bool_:
            a = a.astype(np.int32)          # alps::hdf5 stores bool as an integer
        elif a.dtype.kind == "i":
            a = a.astype("<i8" if a.dtype.itemsize == 8 else "<i4")
        elif a.dtype.kind == "u":
            a = a.astype("<u8" if a.dtype.itemsize == 8 else "<u4")
        elif a.dtype.kind == "f":
            a = a.astype("<f8")
        else:
            raise TypeError("unsupported dtype %s for %s" % (a.dtype, path))
        g.children[parts[-1]] = np.ascontiguousarray(a) if a.ndim else a

    def require_group(self, path):
        g = self.root
        for p in [q for q in path.split("/") if q]:
            g = g.children.setdefault(p, _Group())
        return g

    # ---- serialisation ----
    @staticmethod
    def _message(mtype, data, flags=0):
        data = _pad8(data)
        return struct.pack("<HHB3x", mtype, len(data), flags) + data

    @staticmethod
    def _object_header(messages):
        body = b"".join(messages)
        # version 1, reserved, number of messages, reference count, header data size; 4 bytes of alignment padding
        return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body

    @staticmethod
    def _datatype(value):
        if isinstance(value, bytes):
            # class 3 (string), version 1; null-terminated ASCII
            return struct.pack("<B3BI", 0x13, 0x00, 0, 0, len(value))
        dt = value.dtype
        if dt.kind == "f":
            # class 1 (floating point): little-endian, mantissa normalisation "msb implied", sign at bit 63
            return struct.pack("<B3BI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        signed = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<B3BI", 0x10, signed, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)

    @staticmethod
    def _dataspace(value):
        if isinstance(value, bytes) or value.ndim == 0:
            return struct.pack("<BBB5x", 1, 0, 0)
        return struct.pack("<BBB5x", 1, value.ndim, 0) + b"".join(struct.pack("<Q", d) for d in value.shape)

    def save(self, fname):
        blob = bytearray(96)  # the superblock goes in last

        def alloc(data):
            while len(blob) % 8:
                blob.append(0)
            addr = len(blob)
            blob.extend(data)
            return addr

        def write_dataset(value):
            raw = value if isinstance(value, bytes) else value.tobytes()
            data_addr = alloc(raw) if len(raw) else UNDEF
            msgs = [
                self._message(0x0001, self._dataspace(value)),
                self._message(0x0003, self._datatype(value), flags=1),            # constant message
                self._message(0x0005, struct.pack("<BBBB", 2, 2, 2, 0)),          # fill value v2: late alloc, write if set, undefined
                self._message(0x0008, struct.pack("<BBQQ", 3, 1, data_addr, len(raw))),  # layout v3, contiguous
            ]
            return alloc(self._object_header(msgs))

        def write_group(g):
            names = sorted(g.children, key=lambda s: s.encode())
            if len(names) > 4 * LEAF_K * INTERNAL_K:
                raise ValueError("group with more than %d members" % (4 * LEAF_K * INTERNAL_K))
            entries = []
            for nm in names:
                ch = g.children[nm]
                if isinstance(ch, _Group):
                    oh, bt, hp = write_group(ch)
                    entries.append((nm, oh, 1, bt, hp))
                else:
                    entries.append((nm, write_dataset(ch), 0, 0, 0))
            # local heap: offset 0 holds the empty string, names follow, 8-byte aligned
            heap, offs = bytearray(8), []
            for nm, *_ in entries:
                offs.append(len(heap))
                heap.extend(_pad8(nm.encode() + b"\0"))
            heap_data = alloc(bytes(heap))
            heap_addr = alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 1, heap_data))  # free-list head 1 = none
            # symbol-table nodes (sorted by name, 2 * LEAF_K entries each) under one level-0 B-tree node:
            # key_0 = "", child_i holds the names in (key_i, key_{i+1}], key_{i+1} = largest name of child_i
            keys, kids = [0], []
            for c0 in range(0, len(entries), 2 * LEAF_K):
                chunk = list(zip(entries[c0:c0 + 2 * LEAF_K], offs[c0:c0 + 2 * LEAF_K]))
                snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk)))
                for (nm, oh, cache, bt, hp), off in chunk:
                    snod.extend(struct.pack("<QQII", off, oh, cache, 0) + (struct.pack("<QQ", bt, hp) if cache else bytes(16)))
                snod.extend(bytes(40 * (2 * LEAF_K - len(chunk))))
                kids.append(alloc(bytes(snod)))
                keys.append(chunk[-1][1])
            node = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, len(kids), UNDEF, UNDEF))
            node.extend(struct.pack("<Q", keys[0]))
            for kid, key in zip(kids, keys[1:]):
                node.extend(struct.pack("<QQ", kid, key))
            node.extend(bytes(24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8 - len(node)))
            bt_addr = alloc(bytes(node))
            oh_addr = alloc(self._object_header([self._message(0x0011, struct.pack("<QQ", bt_addr, heap_addr))]))
            return oh_addr, bt_addr, heap_addr

        root_oh, root_bt, root_hp = write_group(self.root)
        while len(blob) % 8:
            blob.append(0)
        sb = SIGNATURE + struct.pack("<8B", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(blob), UNDEF)
        sb += struct.pack("<QQII", 0, root_oh, 1, 0) + struct.pack("<QQ", root_bt, root_hp)
        assert len(sb) == 96
        blob[:96] = sb
        with open(fname, "wb") as fh:
            fh.write(bytes(blob))
        return len(blob)


class H5Reader:
    """Reads the subset H5Writer emits (and that libhdf5 emits with default settings for small files):
    superblock v0 (optionally behind a user block), symbol-table groups, v1 object headers with continuation
    blocks, contiguous or compact datasets of fixed-point / floating-point / fixed-length string type."""

    def __init__(self, fname):
        self.b = open(fname, "rb").read()
        at = 0
        while self.b[at:at + 8] != SIGNATURE:
            at = 512 if at == 0 else at * 2
            if at >= len(self.b):
                raise ValueError("not an HDF5 file")
        sb = self.b[at:at + 96]
        ver, _, _, _, _, so, sl, _ = struct.unpack("<8B", sb[8:16])
        if ver != 0 or so != 8 or sl != 8:
            raise ValueError("unsupported superblock (version %d, offsets %d, lengths %d)" % (ver, so, sl))
        self.leaf_k, self.internal_k, _ = struct.unpack("<HHI", sb[16:24])
        self.base, _, self.eof, _ = struct.unpack("<QQQQ", sb[24:56])
        _, self.root_oh, cache, _, self.root_bt, self.root_hp = struct.unpack("<QQIIQQ", sb[56:96])

    def _at(self, addr, n):
        a = self.base + addr
        return self.b[a:a + n]

    def _messages(self, oh_addr):
        ver, _, nmsg, _, size = struct.unpack("<BBHII", self._at(oh_addr, 12))
        if ver != 1:
            raise ValueError("object header version %d" % ver)
        out, blocks = [], [(oh_addr + 16, size)]
        while blocks and len(out) < nmsg:
            pos, left = blocks.pop(0)
            while left >= 8 and len(out) < nmsg:
                mtype, msize, flags = struct.unpack("<HHB3x", self._at(pos, 8))
                data = self._at(pos + 8, msize)
                if mtype == 0x0010:  # continuation
                    blocks.append(struct.unpack("<QQ", data[:16]))
                out.append((mtype, data))
                pos += 8 + msize
                left -= 8 + msize
        return out

    def _heap_name(self, heap_addr, off):
        sig, _, _, data_addr = self._at(heap_addr, 8)[:4], *struct.unpack("<QQQ", self._at(heap_addr + 8, 24))
        if sig != b"HEAP":
            raise ValueError("bad local heap")
        a = self.base + data_addr + off
        return self.b[a:self.b.index(b"\0", a)].decode()

    def _btree_entries(self, bt_addr, heap_addr):
        hdr = self._at(bt_addr, 24)
        if hdr[:4] != b"TREE":
            raise ValueError("bad B-tree node")
        ntype, level, used = struct.unpack("<BBH", hdr[4:8])
        out = []
        for i in range(used):
            child = struct.unpack("<Q", self._at(bt_addr + 24 + 8 + 16 * i, 8))[0]
            if level > 0:
                out += self._btree_entries(child, heap_addr)
                continue
            sn = self._at(child, 8)
            if sn[:4] != b"SNOD":
                raise ValueError("bad symbol-table node")
            nsym = struct.unpack("<H", sn[6:8])[0]
            for k in range(nsym):
                off, oh, cache, _, s0, s1 = struct.unpack("<QQIIQQ", self._at(child + 8 + 40 * k, 40))
                out.append((self._heap_name(heap_addr, off), oh))
        return out

    def members(self, oh_addr=None):
        """name -> object-header address of a group's members."""
        oh_addr = self.root_oh if oh_addr is None else oh_addr
        for mtype, data in self._messages(oh_addr):
            if mtype == 0x0011:
                bt, hp = struct.unpack("<QQ", data[:16])
                return dict(self._btree_entries(bt, hp))
        return None  # not a group

    def _dataset(self, oh_addr):
        shape = dtype = None
        raw = None
        for mtype, data in self._messages(oh_addr):
            if mtype == 0x0001:
                ver, rank, flags = struct.unpack("<BBB", data[:3])
                o = 8 if ver == 1 else 4
                shape = struct.unpack("<%dQ" % rank, data[o:o + 8 * rank])
            elif mtype == 0x0003:
                cls, b0, b1, b2, size = struct.unpack("<B3BI", data[:8])
                if cls & 15 == 0:
                    dtype = np.dtype(("<" if not b0 & 1 else ">") + ("i" if b0 & 8 else "u") + str(size))
                elif cls & 15 == 1:
                    dtype = np.dtype(("<" if not b0 & 1 else ">") + "f" + str(size))
                elif cls & 15 == 3:
                    dtype = np.dtype("S%d" % size)
                else:
                    raise ValueError("unsupported datatype class %d" % (cls & 15))
            elif mtype == 0x0008:
                ver, b1 = struct.unpack("<BB", data[:2])
                if ver == 3:
                    if b1 == 1:
                        addr, size = struct.unpack("<QQ", data[2:18])
                        raw = ("at", addr)
                    elif b1 == 0:
                        size = struct.unpack("<H", data[2:4])[0]
                        raw = data[4:4 + size]
                    else:
                        raise ValueError("chunked datasets are not supported")
                elif ver in (1, 2):  # version, rank, class, 5 reserved, [address], rank x 4-byte sizes, [compact size + data]
                    rank_l, lclass = b1, data[2]
                    if lclass == 1:
                        raw = ("at", struct.unpack("<Q", data[8:16])[0])
                    elif lclass == 0:
                        o = 8 + 4 * rank_l
                        size = struct.unpack("<I", data[o:o + 4])[0]
                        raw = data[o + 4:o + 4 + size]
                    else:
                        raise ValueError("chunked datasets are not supported")
                else:
                    raise ValueError("layout message version %d" % ver)
        if shape is None or dtype is None or raw is None:
            raise ValueError("not a dataset")
        n = int(np.prod(shape)) if shape else 1
        if isinstance(raw, tuple):
            raw = b"" if raw[1] == UNDEF else self._at(raw[1], n * dtype.itemsize)
        a = np.frombuffer(raw[:n * dtype.itemsize], dtype=dtype).reshape(shape)
        if dtype.kind == "S":
            return a.reshape(-1)[0].split(b"\0")[0].decode() if not shape else a
        return a.copy() if shape else a.reshape(()).item()

    def __getitem__(self, path):
        oh = self.root_oh
        for p in [q for q in path.split("/") if q]:
            m = self.members(oh)
            if m is None or p not in m:
                raise KeyError(path)
            oh = m[p]
        m = self.members(oh)
        return m if m is not None else self._dataset(oh)

    def tree(self, oh_addr=None, prefix=""):
        """Flat {path: value} of everything below a group."""
        out = {}
        for nm, oh in sorted((self.members(oh_addr) or {}).items()):
            if self.members(oh) is not None:
                out.update(self.tree(oh, prefix + "/" + nm))
            else:
                out[prefix + "/" + nm] = self._dataset(oh)
        return out


def save_all_data(fname, params, energies, d2energies, c_energies, beta, volume, max_depth=None, histories=None, nf0=None, nfpi=None,
                  spectrum_history=None, ipr_history=None, dos_wgrid=None, dos_offset=0.05):
    """Reference layout of prog/data_save.hxx (save_all_data -> save_measurements + energy / cv statistics).

    params: dict of run parameters (the alps::params dump); energies, d2energies, c_energies: 1-D series (one chain, or
    already pooled) or 2-D [measurement][chain] as fkmc_chain_get_series returns them -- pooled CHAIN-MAJOR like the
    reference's rank-by-rank gather (stats.pool_chains); histories: optional dict of 2-D
    [index][measurement] arrays for /mc_data (ipr_history, spectrum_history, focc_history).
    nf0 / nfpi: optional f-sector series (fkmc_chain_get_fsector) -> /mc_data/{nf0,nfpi} and the save_fstats statistics
    (nf_0, nf_pi, fsusc_0, fsusc_pi, binder_0, binder_pi; prog/data_save.hxx:200-236).
    spectrum_history / ipr_history ([measurement][chain][N] as fkmc_chain_get_history returns them) with dos_wgrid: the DOS and IPR
    post-processing of save_glocal / save_ipr (prog/data_save.hxx:265-345,487-532): dos0, dos_err, nc, ipr0, ipr_err; the histories
    themselves go to /mc_data/{spectrum_history,ipr_history} as [index][measurement] (prog/data_save.hxx:105-137).
    Returns the per-observable statistics that were written (dict name -> (binning rows, stats 4-vector))."""
    w = H5Writer()
    w.require_group("/parameters")
    for k, v in sorted(params.items()):
        w["/parameters/" + k] = v
    e = stats.pool_chains(energies)
    d2 = stats.pool_chains(d2energies)
    ce = stats.pool_chains(c_energies)
    w["/mc_data/energies"] = e
    w["/mc_data/d2energies"] = d2
    w["/mc_data/c_energies"] = ce
    for k, v in (histories or {}).items():
        w["/mc_data/" + k] = np.asarray(v, dtype=np.float64)
    if max_depth is None:
        max_depth = stats.max_bin_depth(e.size)
    rep = stats.energy_report(e, d2, beta, volume, max_depth)
    out = {}

    def put(name, rows):
        # /binning/<obs>: nbins x 5 [n, mean, variance, stderr, tau_int]; /stats/<obs>: the row estimate_bin picks (data_save.hpp:124-152)
        table = np.array([list(r) + [c] for r, c in zip(rows, stats.calc_cor_length(rows))], dtype=np.float64)
        st = np.array(rows[stats.estimate_bin(rows)], dtype=np.float64)
        w["/binning/" + name] = table
        w["/stats/" + name] = st
        out[name] = (table, st)

    put("energy", rep["energy"]["binning"])
    put("d2energy", rep["d2energy"]["binning"])
    put("c_energy", stats.accumulate_binning(ce[::-1], max_depth))  # the reference bins the reversed series
    put("cv", rep["cv"]["binning"])
    if nf0 is not None and nfpi is not None:
        w["/mc_data/nf0"] = stats.pool_chains(nf0)
        w["/mc_data/nfpi"] = stats.pool_chains(nfpi)
        frep = stats.fstats_report(nf0, nfpi, max_depth)
        for name in ("nf_0", "nf_pi", "fsusc_0", "fsusc_pi"):
            put(name, frep[name]["binning"])
        for name in ("binder_0", "binder_pi"):   # save_bin_data: /stats only
            st = np.array(frep[name]["stats"], dtype=np.float64)
            w["/stats/" + name] = st
            out[name] = (None, st)
    if spectrum_history is not None:
        w["/mc_data/spectrum_history"] = np.ascontiguousarray(stats._history_rows(spectrum_history).T)   # [index][measurement]
        if dos_wgrid is not None:
            drep = stats.dos_report(spectrum_history, dos_wgrid, dos_offset, beta, max_depth)
            put("dos0", drep["dos0"]["binning"])
            w["/stats/dos_err"] = drep["dos_err"]
            if "nc" in drep:
                w["/stats/nc"] = np.array(drep["nc"], dtype=np.float64)
    if ipr_history is not None:
        w["/mc_data/ipr_history"] = np.ascontiguousarray(stats._history_rows(ipr_history).T)
        if spectrum_history is not None and dos_wgrid is not None:
            irep = stats.ipr_report(spectrum_history, ipr_history, dos_wgrid, dos_offset, max_depth)
            put("ipr0", irep["ipr0"]["binning"])
            w["/stats/ipr_err"] = irep["ipr_err"]
    w.save(fname)
    return out
