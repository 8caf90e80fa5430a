"""On-disk twins of the reference's matrix files.

The reference stores its results with ibmisc's NcIO (netCDF-4): `linear::Weighted_Eigen::ncio` writes the "Eigen format"
(sphinx/source/matrix_formats.rst:9-63; call sites modele/global_ec.cpp:566-659), `combine_global_ec` the "compressed
format" (matrix_formats.rst:65-117, modele/combine_global_ec.cpp:95-262) and `GCMRegridder_Standard::ncio` the regridder
itself (slib/icebin/GCMRegridder.cpp:104-150, AbbrGrid.cpp:23-28, 167-191, IceRegridder.cpp:75-89).  Neither libnetcdf nor
HDF5 exists in this image, so a `Dataset` here is a plain in-memory model of a netCDF file -- dimensions, variables,
attributes, with the reference's names, types and shapes -- and three serialisations:

  * CDL text (`to_cdl` / `from_cdl`): the netCDF-4 CDL that `ncdump` prints and `ncgen -k nc4` turns into the very file the
    reference writes (int64 variables, `string` attributes).  This is the interoperable form.
  * CDF-5 binary (`to_cdf5` / `from_cdf5`): the classic netCDF format with 64-bit types, readable by libnetcdf >= 4.4
    directly; `string` attributes become one char attribute with the names separated by blanks.
  * NPZ (`to_npz` / `from_npz`): the same names as numpy arrays, for speed.

Matrices travel as host arrays: `HostWeighted` is what `Weighted_Eigen` holds (dense-indexed COO + weights + the two
SparseSets); `to_device()` hands it to the GPU library, `from_device()` takes it back.
"""
import io
import json
import re
import struct
import zlib
from collections import OrderedDict

import numpy as np

# ---------------------------------------------------------------------------------------------------- dataset model
_CDL_TYPES = OrderedDict([("byte", np.int8), ("ubyte", np.uint8), ("short", np.int16), ("int", np.int32), ("int64", np.int64),
                          ("float", np.float32), ("double", np.float64), ("char", "S1")])
_NC_CODE = {np.dtype(np.int8): 1, np.dtype("S1"): 2, np.dtype(np.int16): 3, np.dtype(np.int32): 4, np.dtype(np.float32): 5,
            np.dtype(np.float64): 6, np.dtype(np.uint8): 7, np.dtype(np.uint16): 8, np.dtype(np.uint32): 9, np.dtype(np.int64): 10,
            np.dtype(np.uint64): 11}
_NC_DTYPE = {v: k for k, v in _NC_CODE.items()}


def _cdl_type(dt):
    dt = np.dtype(dt)
    for name, t in _CDL_TYPES.items():
        if np.dtype(t) == dt:
            return name
    raise TypeError("no CDL type for %s" % dt)


class Variable:
    def __init__(self, dtype, dims, data, attrs=None):
        self.dtype = np.dtype(dtype)
        self.dims = tuple(dims)
        self.data = np.ascontiguousarray(data, dtype=self.dtype)
        self.attrs = OrderedDict(attrs or {})


class Dataset:
    """dims: name -> length; vars: name -> Variable; attrs: global attributes.  Attribute values: str, list of str
    (netCDF-4 `string` array), or a numpy scalar / 1-D array (the dtype is the attribute's type)."""

    def __init__(self, name="icebin"):
        self.name = name
        self.dims = OrderedDict()
        self.vars = OrderedDict()
        self.attrs = OrderedDict()

    # -- building blocks (ibmisc get_or_add_dim / get_or_add_var / ncio_vector)
    def add_dim(self, name, n):
        n = int(n)
        if name in self.dims and self.dims[name] != n:
            raise ValueError("dimension %s redefined: %d vs %d" % (name, self.dims[name], n))
        self.dims[name] = n
        return name

    def add_var(self, name, dtype, dims, data, attrs=None):
        data = np.asarray(data)
        shape = tuple(self.dims[d] for d in dims)
        if data.shape != shape:
            raise ValueError("variable %s: data shape %s does not match dimensions %s" % (name, data.shape, shape))
        self.vars[name] = Variable(dtype, dims, data, attrs)
        return self.vars[name]

    def info_var(self, name, dtype=np.int32, attrs=None):
        """A scalar variable that only carries attributes (get_or_add_var(ncio, vname + ".info", "int", {}))."""
        self.vars[name] = Variable(dtype, (), np.zeros((), dtype), attrs)
        return self.vars[name]

    # -- CDL
    def to_cdl(self, path=None):
        out = io.StringIO()
        out.write("netcdf %s {\n" % self.name)
        if self.dims:
            out.write("dimensions:\n")
            for k, n in self.dims.items():
                out.write("\t%s = %d ;\n" % (k, n))
        out.write("variables:\n")
        for k, v in self.vars.items():
            dims = "(%s)" % ", ".join(v.dims) if v.dims else ""
            out.write("\t%s %s%s ;\n" % (_cdl_type(v.dtype), k, dims))
            for an, av in v.attrs.items():
                out.write("\t\t%s ;\n" % _cdl_att(k, an, av))
        if self.attrs:
            out.write("\n// global attributes:\n")
            for an, av in self.attrs.items():
                out.write("\t\t%s ;\n" % _cdl_att("", an, av))
        out.write("data:\n")
        for k, v in self.vars.items():
            if v.data.size == 0 or not v.dims:
                continue
            out.write("\n %s = " % k)
            flat = v.data.reshape(-1)
            out.write(", ".join(_cdl_values(flat)))
            out.write(" ;\n")
        out.write("}\n")
        text = out.getvalue()
        if path:
            with open(path, "w") as f:
                f.write(text)
        return text

    @classmethod
    def from_cdl(cls, text_or_path):
        text = text_or_path
        if "\n" not in text_or_path and not text_or_path.lstrip().startswith("netcdf"):
            with open(text_or_path) as f:
                text = f.read()
        m = re.match(r"\s*netcdf\s+(\S+)\s*\{", text)
        ds = cls(m.group(1) if m else "icebin")
        body = text[m.end():] if m else text
        sec = re.split(r"\n\s*(dimensions|variables|data):\s*\n", "\n" + body)
        parts = {sec[i]: sec[i + 1] for i in range(1, len(sec) - 1, 2)}
        for stmt in _cdl_statements(parts.get("dimensions", "")):
            k, n = [t.strip() for t in stmt.split("=")]
            ds.dims[k] = int(n)
        decl = re.compile(r"^(%s)\s+([^\s(]+)\s*(?:\(([^)]*)\))?$" % "|".join(_CDL_TYPES))
        for stmt in _cdl_statements(parts.get("variables", "")):
            dm = decl.match(stmt)
            if dm and "=" not in stmt.split("(")[0]:
                dims = tuple(d.strip() for d in dm.group(3).split(",")) if dm.group(3) else ()
                shape = tuple(ds.dims[d] for d in dims)
                ds.vars[dm.group(2)] = Variable(_CDL_TYPES[dm.group(1)], dims, np.zeros(shape, _CDL_TYPES[dm.group(1)]))
                continue
            lhs, rhs = stmt.split("=", 1)
            lhs = lhs.strip()
            typ = None
            tm = re.match(r"^(string|%s)\s+(.*)$" % "|".join(_CDL_TYPES), lhs)
            if tm:
                typ, lhs = tm.group(1), tm.group(2).strip()
            var, att = lhs.rsplit(":", 1)
            val = _cdl_parse_att(rhs.strip(), typ)
            (ds.vars[var].attrs if var else ds.attrs)[att] = val
        for stmt in _cdl_statements(parts.get("data", "").rsplit("}", 1)[0]):
            k, rhs = stmt.split("=", 1)
            v = ds.vars[k.strip()]
            toks = [t.strip() for t in rhs.split(",") if t.strip()]
            v.data = np.array([_cdl_number(t, v.dtype) for t in toks], dtype=v.dtype).reshape(v.data.shape)
        return ds

    # -- NPZ
    def to_npz(self, path):
        meta = {"name": self.name, "dims": list(self.dims.items()),
                "vars": [(k, _cdl_type(v.dtype), list(v.dims), _json_attrs(v.attrs)) for k, v in self.vars.items()],
                "attrs": _json_attrs(self.attrs)}
        arrays = {"var:" + k: v.data for k, v in self.vars.items()}
        np.savez(path, __meta__=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)

    @classmethod
    def from_npz(cls, path):
        z = np.load(path, allow_pickle=False)
        meta = json.loads(bytes(z["__meta__"]).decode())
        ds = cls(meta["name"])
        ds.dims = OrderedDict((k, int(n)) for k, n in meta["dims"])
        for k, typ, dims, attrs in meta["vars"]:
            ds.vars[k] = Variable(_CDL_TYPES[typ], dims, z["var:" + k], _unjson_attrs(attrs))
        ds.attrs = _unjson_attrs(meta["attrs"])
        return ds

    # -- CDF-5 (classic netCDF with 64-bit types; big-endian, every variable padded to a 4-byte boundary)
    def to_cdf5(self, path):
        dimid = {k: i for i, k in enumerate(self.dims)}

        def name(s):
            b = s.encode()
            return struct.pack(">q", len(b)) + b + b"\0" * (-len(b) % 4)

        def att_list(attrs):
            if not attrs:
                return struct.pack(">iq", 0, 0)
            out = struct.pack(">iq", 12, len(attrs))
            for an, av in attrs.items():
                if isinstance(av, (list, tuple)) and all(isinstance(x, str) for x in av):
                    av = " ".join(av)                       # no NC_STRING in the classic model
                if isinstance(av, str):
                    b = av.encode()
                    out += name(an) + struct.pack(">iq", 2, len(b)) + b + b"\0" * (-len(b) % 4)
                else:
                    a = np.atleast_1d(np.asarray(av))
                    b = a.astype(a.dtype.newbyteorder(">")).tobytes()
                    out += name(an) + struct.pack(">iq", _NC_CODE[a.dtype], a.size) + b + b"\0" * (-len(b) % 4)
            return out

        def header(begins):
            h = b"CDF\x05" + struct.pack(">q", 0)
            h += struct.pack(">iq", 10, len(self.dims)) if self.dims else struct.pack(">iq", 0, 0)
            for k, n in self.dims.items():
                h += name(k) + struct.pack(">q", n)
            h += att_list(self.attrs)
            h += struct.pack(">iq", 11, len(self.vars)) if self.vars else struct.pack(">iq", 0, 0)
            for (k, v), beg in zip(self.vars.items(), begins):
                h += name(k) + struct.pack(">q", len(v.dims)) + b"".join(struct.pack(">q", dimid[d]) for d in v.dims)
                nbytes = v.data.size * v.dtype.itemsize
                h += att_list(v.attrs) + struct.pack(">iqq", _NC_CODE[v.dtype], nbytes + (-nbytes % 4), beg)
            return h

        hlen = len(header([0] * len(self.vars)))
        begins, pos = [], hlen
        for v in self.vars.values():
            begins.append(pos)
            nbytes = v.data.size * v.dtype.itemsize
            pos += nbytes + (-nbytes % 4)
        with open(path, "wb") as f:
            f.write(header(begins))
            for v in self.vars.values():
                b = v.data.astype(v.dtype.newbyteorder(">") if v.dtype.kind != "S" else v.dtype).tobytes()
                f.write(b + b"\0" * (-len(b) % 4))

    @classmethod
    def from_cdf5(cls, path):
        with open(path, "rb") as f:
            buf = f.read()
        if buf[:4] != b"CDF\x05":
            raise ValueError("not a CDF-5 file")
        pos = [4]

        def rd(fmt):
            vals = struct.unpack_from(">" + fmt, buf, pos[0])
            pos[0] += struct.calcsize(">" + fmt)
            return vals if len(vals) > 1 else vals[0]

        def name():
            n = rd("q")
            s = buf[pos[0]:pos[0] + n].decode()
            pos[0] += n + (-n % 4)
            return s

        def att_list():
            tag, n = rd("iq")
            attrs = OrderedDict()
            for _ in range(n if tag == 12 else 0):
                an = name()
                typ, ne = rd("iq")
                dt = _NC_DTYPE[typ]
                nb = ne * dt.itemsize
                raw = buf[pos[0]:pos[0] + nb]
                pos[0] += nb + (-nb % 4)
                if typ == 2:
                    attrs[an] = raw.decode()
                else:
                    a = np.frombuffer(raw, dtype=dt.newbyteorder(">")).astype(dt)
                    attrs[an] = a[0] if ne == 1 else a
            return attrs

        rd("q")                                             # numrecs
        ds = cls()
        tag, n = rd("iq")
        for _ in range(n if tag == 10 else 0):
            k = name()
            ds.dims[k] = rd("q")
        ds.attrs = att_list()
        tag, n = rd("iq")
        dimnames = list(ds.dims)
        for _ in range(n if tag == 11 else 0):
            k = name()
            nd = rd("q")
            dims = tuple(dimnames[rd("q")] for _ in range(nd))
            attrs = att_list()
            typ, vsize, begin = rd("iqq")
            dt = _NC_DTYPE[typ]
            shape = tuple(ds.dims[d] for d in dims)
            cnt = int(np.prod(shape)) if shape else 1
            data = np.frombuffer(buf, dtype=dt.newbyteorder(">") if dt.kind != "S" else dt, count=cnt, offset=begin).astype(dt).reshape(shape)
            ds.vars[k] = Variable(dt, dims, data, attrs)
        return ds


def _cdl_statements(block):
    """Statements of a CDL section: text up to ' ;' outside quotes, comments removed."""
    out, cur, inq = [], [], False
    for line in block.splitlines():
        if not inq and line.strip().startswith("//"):
            continue
        i = 0
        while i < len(line):
            ch = line[i]
            if ch == '"' and (i == 0 or line[i - 1] != "\\"):
                inq = not inq
            if ch == ";" and not inq:
                s = "".join(cur).strip()
                if s:
                    out.append(s)
                cur = []
            else:
                cur.append(ch)
            i += 1
        cur.append(" ")
    return out


def _cdl_values(flat):
    if flat.dtype.kind == "f":
        return [("NaN" if np.isnan(x) else ("Infinity" if x > 0 else "-Infinity") if np.isinf(x) else repr(float(x))) for x in flat]
    if flat.dtype == np.int64:
        return ["%d" % x for x in flat]
    return ["%d" % x for x in flat]


def _cdl_att(var, name, val):
    if isinstance(val, str):
        return '%s:%s = "%s"' % (var, name, val.replace("\\", "\\\\").replace('"', '\\"'))
    if isinstance(val, (list, tuple)) and all(isinstance(x, str) for x in val):
        return 'string %s:%s = %s' % (var, name, ", ".join('"%s"' % x for x in val))
    a = np.atleast_1d(np.asarray(val))
    suffix = {"int64": "LL", "byte": "b", "short": "s", "float": "f", "ubyte": "UB"}.get(_cdl_type(a.dtype), "")
    if a.dtype.kind == "f":
        vals = [repr(float(x)) + ("f" if a.dtype == np.float32 else "") for x in a]
    else:
        vals = ["%d%s" % (x, suffix) for x in a]
    return "%s:%s = %s" % (var, name, ", ".join(vals))


def _cdl_number(tok, dtype):
    if tok in ("NaN", "NaNf", "_"):
        return np.nan
    if tok in ("Infinity", "-Infinity"):
        return np.inf if tok[0] != "-" else -np.inf
    tok = re.sub(r"(LL|ULL|UB|b|s|L|U)$", "", tok)
    if np.dtype(dtype).kind == "f":
        return float(tok[:-1] if tok.endswith("f") else tok)
    return int(tok)


def _cdl_parse_att(rhs, typ):
    if rhs.startswith('"'):
        strs = [s.replace('\\"', '"').replace("\\\\", "\\") for s in re.findall(r'"((?:[^"\\]|\\.)*)"', rhs)]
        return strs if typ == "string" else strs[0]
    toks = [t.strip() for t in rhs.split(",")]
    if all(re.match(r"^-?\d+(LL|L)$", t) for t in toks):
        a = np.array([int(re.sub(r"L+$", "", t)) for t in toks], np.int64)
    elif all(re.match(r"^-?\d+$", t) for t in toks):
        a = np.array([int(t) for t in toks], np.int32)
    elif all(t.endswith("f") and not t.startswith("0x") for t in toks):
        a = np.array([float(t[:-1]) for t in toks], np.float32)
    else:
        a = np.array([_cdl_number(t, np.float64) for t in toks], np.float64)
    return a[0] if a.size == 1 else a


def _json_attrs(attrs):
    out = []
    for k, v in attrs.items():
        if isinstance(v, str):
            out.append([k, "str", v])
        elif isinstance(v, (list, tuple)) and all(isinstance(x, str) for x in v):
            out.append([k, "strings", list(v)])
        else:
            a = np.atleast_1d(np.asarray(v))
            out.append([k, _cdl_type(a.dtype) + ("" if np.ndim(v) else ":scalar"), [float(x) if a.dtype.kind == "f" else int(x) for x in a]])
    return out


def _unjson_attrs(lst):
    out = OrderedDict()
    for k, typ, v in lst:
        if typ == "str":
            out[k] = v
        elif typ == "strings":
            out[k] = list(v)
        else:
            scalar = typ.endswith(":scalar")
            a = np.array(v, dtype=_CDL_TYPES[typ.split(":")[0]])
            out[k] = a[0] if scalar else a
    return out


# ---------------------------------------------------------------------------------------------------- Weighted_Eigen
class HostWeighted:
    """What linear::Weighted_Eigen holds, on the host: a dense-indexed COO matrix, its weights, and the two SparseSets
    (dense -> sparse ids and sparse extents)."""

    def __init__(self, shape, row, col, val, wM, Mw, dims, sparse_extents, conservative=True):
        self.shape = (int(shape[0]), int(shape[1]))
        self.row, self.col = np.asarray(row, np.int32), np.asarray(col, np.int32)
        self.val = np.asarray(val, np.float64)
        self.wM, self.Mw = np.asarray(wM, np.float64), np.asarray(Mw, np.float64)
        self.dims = [np.asarray(dims[0], np.int64), np.asarray(dims[1], np.int64)]
        self.sparse_extents = [int(sparse_extents[0]), int(sparse_extents[1])]
        self.conservative = bool(conservative)

    @classmethod
    def from_device(cls, W):
        """W: icebin_b200.api.Weighted.  Entries in the reference's to_coo() order."""
        r, c, v = W.to_coo_dense()
        return cls(W.shape, r, c, v, W.get_weights(0), W.get_weights(1), [W.dim(0), W.dim(1)], W.sparse_extent, W.conservative)

    def to_device(self):
        """A GPU-resident icebin_b200.api.Weighted with identity dims over the DENSE spaces (ib_w_create_csr); use
        `dims` / `sparse_extents` to map to sparse ids, or apply_M_host below."""
        from . import api
        o = np.lexsort((self.col, self.row))
        rowptr = np.zeros(self.shape[0] + 1, np.int32)
        np.add.at(rowptr, self.row + 1, 1)
        rowptr = np.cumsum(rowptr).astype(np.int32)
        return api.Weighted.from_csr(self.shape[0], self.shape[1], rowptr, self.col[o], self.val[o], self.wM, self.Mw, self.conservative)

    def to_coo(self):
        """matrix_formats.rst 'Extraction': scipy COO in SPARSE indices."""
        import scipy.sparse
        return scipy.sparse.coo_matrix((self.val, (self.dims[0][self.row], self.dims[1][self.col])), shape=tuple(self.sparse_extents))


def put_dim(ds, name, ids, sparse_extent, shape=None, description=None):
    """SparseSetT::ncio: `int64 dimA(dimA.dense_extent)` with attribute sparse_extent (matrix_formats.rst:19-31); shape and
    description as modele/global_ec.cpp:631-652 adds them."""
    ids = np.asarray(ids, np.int64)
    ds.add_dim(name + ".dense_extent", len(ids))
    v = ds.add_var(name, np.int64, (name + ".dense_extent",), ids, OrderedDict(sparse_extent=np.int64(sparse_extent)))
    if shape is not None:
        v.attrs["shape"] = np.asarray(shape, np.int32)
    if description is not None:
        v.attrs["description"] = description
    return v


def put_weighted(ds, W, vname, dim_names):
    """Weighted_Eigen::ncio(ncio, vname, {dimB_name, dimA_name}) (matrix_formats.rst:17-42; mat->ncio(ncio, "EvI",
    {"dimE", "dimI"}), modele/global_ec.cpp:581).  The dims themselves are written by put_dim (global_ec.cpp:628-652)."""
    nnz = len(W.val)
    for k, dn in enumerate(dim_names):
        ds.add_dim(dn + ".dense_extent", W.shape[k])
    ds.add_dim(vname + ".M.nnz", nnz)
    ds.add_dim(vname + ".M.rank", 2)
    ds.info_var(vname + ".info", np.int32, OrderedDict(type="EIGEN", conservative=np.int32(W.conservative), dim_names=list(dim_names)))
    ds.info_var(vname + ".M.info", np.int64, OrderedDict(shape=np.asarray(W.shape, np.int64), conservative="t" if W.conservative else "f"))
    ds.add_var(vname + ".M.indices", np.int32, (vname + ".M.nnz", vname + ".M.rank"), np.stack([W.row, W.col], axis=1).reshape(nnz, 2))
    ds.add_var(vname + ".M.values", np.float64, (vname + ".M.nnz",), W.val)
    ds.add_var(vname + ".Mw", np.float64, (dim_names[1] + ".dense_extent",), W.Mw)
    ds.add_var(vname + ".wM", np.float64, (dim_names[0] + ".dense_extent",), W.wM)


def get_weighted(ds, vname):
    info = ds.vars[vname + ".info"].attrs
    if info.get("type") != "EIGEN":
        raise ValueError("%s is not stored in Eigen format (type = %r)" % (vname, info.get("type")))
    names = info["dim_names"]
    if isinstance(names, str):
        names = names.split()
    names = [n.split(".")[-1] if n not in ds.vars else n for n in names]      # the doc's example writes "BvA.dimB"
    idx = ds.vars[vname + ".M.indices"].data
    shape = np.atleast_1d(ds.vars[vname + ".M.info"].attrs["shape"])
    dims = [ds.vars[n].data for n in names]
    ext = [int(ds.vars[n].attrs["sparse_extent"]) for n in names]
    return HostWeighted(shape, idx[:, 0], idx[:, 1], ds.vars[vname + ".M.values"].data, ds.vars[vname + ".wM"].data,
                        ds.vars[vname + ".Mw"].data, dims, ext, bool(int(info["conservative"])))


# ---------------------------------------------------------------------------------------------------- compressed format
def _zarray(ds, vname, shape, indices, values):
    """A compressed tensor of rank len(shape) (matrix_formats.rst:85-104): coordinate indices [nnz, rank] and values [nnz],
    each a zlib stream.  The byte layout INSIDE the streams is ibmisc's ZArray (not in the reference tree); here it is the
    little-endian int64 index array in C order and the little-endian fp64 values -- parity unpinned."""
    indices = np.ascontiguousarray(np.asarray(indices, "<i8").reshape(len(values), len(shape)))
    zi = np.frombuffer(zlib.compress(indices.tobytes(), 6), np.uint8)
    zv = np.frombuffer(zlib.compress(np.asarray(values, "<f8").tobytes(), 6), np.uint8)
    ds.info_var(vname + ".info", np.int32, OrderedDict(format="ZARRAY", rank=np.int32(len(shape)), nnz=np.int64(len(values)),
                                                       shape=np.asarray(shape, np.int64)))
    ds.add_dim(vname + ".indices.zsize", len(zi))
    ds.add_dim(vname + ".values.zsize", len(zv))
    ds.add_var(vname + ".indices", np.uint8, (vname + ".indices.zsize",), zi)
    ds.add_var(vname + ".values", np.uint8, (vname + ".values.zsize",), zv)


def _unzarray(ds, vname):
    info = ds.vars[vname + ".info"].attrs
    rank, nnz = int(info["rank"]), int(info["nnz"])
    idx = np.frombuffer(zlib.decompress(ds.vars[vname + ".indices"].data.tobytes()), "<i8").reshape(nnz, rank)
    val = np.frombuffer(zlib.decompress(ds.vars[vname + ".values"].data.tobytes()), "<f8")
    return tuple(int(x) for x in np.atleast_1d(info["shape"])), idx.astype(np.int64), val.astype(np.float64)


def put_compressed(ds, vname, shape, M_row, M_col, M_val, wM_index, wM_val, Mw_index, Mw_val, conservative=True):
    """linear::Weighted_Compressed::ncio as combine_global_ec writes it (modele/combine_global_ec.cpp:256-261): matrix and
    weights in the ORIGINAL sparse spaces, as compressed tensors (matrix_formats.rst:65-117)."""
    ds.info_var(vname + ".info", np.int32, OrderedDict(type="COMPRESSED", conservative=np.int32(conservative)))
    _zarray(ds, vname + ".wM", (shape[0],), np.asarray(wM_index).reshape(-1, 1), wM_val)
    _zarray(ds, vname + ".M", shape, np.stack([M_row, M_col], axis=1), M_val)
    _zarray(ds, vname + ".Mw", (shape[1],), np.asarray(Mw_index).reshape(-1, 1), Mw_val)


def get_compressed(ds, vname):
    """-> dict(shape, M=(row, col, val), wM=(index, val), Mw=(index, val), conservative)"""
    info = ds.vars[vname + ".info"].attrs
    if info.get("type") != "COMPRESSED":
        raise ValueError("%s is not stored in compressed format" % vname)
    shape, mi, mv = _unzarray(ds, vname + ".M")
    _, wi, wv = _unzarray(ds, vname + ".wM")
    _, ci, cv = _unzarray(ds, vname + ".Mw")
    return dict(shape=shape, M=(mi[:, 0], mi[:, 1], mv), wM=(wi[:, 0], wv), Mw=(ci[:, 0], cv), conservative=bool(int(info["conservative"])))


# ---------------------------------------------------------------------------------------------------- global_ec output
def write_global_ec(ds, rm, matrix_names, dims, params, shapes=None, Achar="A"):
    """The matrix section of modele/global_ec.cpp:566-659: every requested matrix through matrix_d with the SHARED dims
    (dimA, dimE, dimI grow as the matrices are generated), stored in Eigen format under the reference's names, then the
    dims with their shape / description attributes.  rm: api.RegridMatrices; dims: {'A': SparseSet, 'E': ..., 'I': ...};
    params: dict(scale=, correctA=, sigma=)."""
    desc = {"A": "GCM ('Atmosphere' or 'Ocean') Grid", "E": "Elevation Grid", "I": "Fine-scale ('Ice') Grid"}
    for spec in ("AvI", "EvI", "IvE", "IvA", "AvE", "EvA"):          # the reference's generation order
        if spec not in matrix_names:
            continue
        b, a = spec[0], spec[2]
        W = rm.matrix_d(spec, (dims[b], dims[a]), **params)
        name = spec.replace("A", Achar)
        put_weighted(ds, HostWeighted.from_device(W), name, ["dim" + (Achar if b == "A" else b), "dim" + (Achar if a == "A" else a)])
    for k, d in dims.items():
        nm = "dim" + (Achar if k == "A" else k)
        # matrices written before a dim grew carry its earlier, shorter extent only through their own nnz: the dim written
        # here is the final one, exactly as upstream stores the dims last (global_ec.cpp:628)
        ds.dims.pop(nm + ".dense_extent", None)
        put_dim(ds, nm, d.ids(), d.sparse_extent, None if shapes is None else shapes.get(k), desc.get(k))
    return ds


# ---------------------------------------------------------------------------------------------------- GCMRegridder::ncio
def put_gcm(ds, gridA, hcdefs, sheets, correctA=True, vname="m"):
    """GCMRegridder_Standard::ncio (GCMRegridder.cpp:104-150) + AbbrGrid::ncio (AbbrGrid.cpp:167-191) + IceRegridder::ncio
    (IceRegridder.cpp:75-89) + ExchangeGrid::ncio (AbbrGrid.cpp:23-28), for what the hot path reads of them.
    gridA: synth.GridA; sheets: list of synth.SheetGrid."""
    ds.info_var(vname + ".info", np.int32, OrderedDict(correctA=np.int32(correctA), sheets=[s.name for s in sheets]))
    ds.add_dim(vname + ".nhc", len(hcdefs))
    ds.add_var(vname + ".hcdefs", np.float64, (vname + ".nhc",), hcdefs)
    _put_abbr(ds, vname + ".agridA", gridA.dim, gridA.nA, gridA.native_area, None, "A")
    for s in sheets:
        vn = vname + "." + s.name
        ds.info_var(vn + ".info", np.int32, OrderedDict(name=s.name, interp_style="ELEV_CLASS_INTERP" if s.interp_style else "Z_INTERP"))
        ds.add_dim("agridA.ndata", len(s.gridA_proj_area))
        ds.add_var(vn + ".gridA_proj_area", np.float64, ("agridA.ndata",), s.gridA_proj_area)
        _put_abbr(ds, vn + ".agridI", s.dimI, s.nI, s.native_areaI, s.centroid_xy, s.name)
        ds.add_dim(vn + ".aexgrid.nindices", 2 * s.nX)
        ds.add_dim(vn + ".aexgrid.noverlaps", s.nX)
        ds.add_var(vn + ".aexgrid.indices", np.int32, (vn + ".aexgrid.nindices",), np.asarray(s.ex_ij, np.int32).reshape(-1))
        ds.add_var(vn + ".aexgrid.overlaps", np.float64, (vn + ".aexgrid.noverlaps",), s.ex_area)
    return ds


def _put_abbr(ds, vn, dim, sparse_extent, native_area, centroid_xy, name):
    ds.info_var(vn + ".info", np.int32, OrderedDict(name=name, coordinates="XY" if centroid_xy is not None else "LONLAT", parameterization="L0"))
    put_dim(ds, vn + ".dim", dim, sparse_extent)
    ds.add_var(vn + ".native_area", np.float64, (vn + ".dim.dense_extent",), native_area)
    if centroid_xy is not None:
        ds.add_dim("two", 2)
        ds.add_var(vn + ".centroid_xy", np.float64, (vn + ".dim.dense_extent", "two"), centroid_xy)


def get_gcm(ds, vname="m"):
    """-> (GridA, hcdefs, [SheetGrid], correctA) as icebin_b200.synth dataclasses (what api.GCMRegridder / add_sheet take)."""
    from . import synth
    info = ds.vars[vname + ".info"].attrs
    names = info["sheets"]
    if isinstance(names, str):
        names = names.split()
    dA = ds.vars[vname + ".agridA.dim"]
    gridA = synth.GridA(0, 0, int(dA.attrs["sparse_extent"]), dA.data.copy(), ds.vars[vname + ".agridA.native_area"].data.copy())
    sheets = []
    for nm in names:
        vn = vname + "." + nm
        dI = ds.vars[vn + ".agridI.dim"]
        cen = ds.vars.get(vn + ".agridI.centroid_xy")
        ex = ds.vars[vn + ".aexgrid.indices"].data.reshape(-1, 2)
        style = ds.vars[vn + ".info"].attrs.get("interp_style", "Z_INTERP")
        sheets.append(synth.SheetGrid(nm, 0, 0, 0.0, int(dI.attrs["sparse_extent"]), dI.data.copy(), ds.vars[vn + ".agridI.native_area"].data.copy(),
                                      None if cen is None else cen.data.copy(), ex.astype(np.int32), ds.vars[vn + ".aexgrid.overlaps"].data.copy(),
                                      ds.vars[vn + ".gridA_proj_area"].data.copy(), 1 if style == "ELEV_CLASS_INTERP" else 0))
    return gridA, ds.vars[vname + ".hcdefs"].data.copy(), sheets, bool(int(info["correctA"]))
