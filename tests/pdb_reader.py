"""Independent minimal reader of PDB (format II) files, enough for the Silo files host/silo_pdb.c
writes: header and primitive formats, structure chart, symbol table, extras, array variables, and
struct variables with pointers (the "itag" encoding), hence Silo's Group objects.

It is driven by what the file declares (type sizes, struct members, alignments), not by knowledge of
the writer, so a layout mistake in the writer shows up as a parse error or wrong values here.
"""
import re
import struct

import numpy as np


class PDBError(ValueError):
    pass


class PDBFile:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if not b.startswith(b"!<<PDB:II>>!\n"):
            raise PDBError("not a PDB II file")
        pos = 13
        n = b[pos]
        fmt = b[pos + 1:pos + n]
        pos += n
        self.sizes = dict(zip(("*", "short", "int", "long", "float", "double"), fmt[0:6]))
        orders = fmt[6:9]
        if any(o != 2 for o in orders):
            raise PDBError("only little-endian files are supported by this reader")
        forder, dorder = fmt[9:9 + self.sizes["float"]], fmt[9 + self.sizes["float"]:9 + self.sizes["float"] + self.sizes["double"]]
        if list(forder) != [4, 3, 2, 1] or list(dorder) != [8, 7, 6, 5, 4, 3, 2, 1]:
            raise PDBError("unexpected floating point byte order")
        rest = fmt[9 + self.sizes["float"] + self.sizes["double"]:]
        self.float_format, self.double_format = list(rest[:7]), list(rest[7:14])
        if self.float_format != [32, 8, 23, 0, 1, 9, 0] or self.double_format != [64, 11, 52, 0, 1, 12, 0]:
            raise PDBError("not IEEE-754 formats")
        line, pos = self._line(pos)
        biases = line.split(b"\x01")
        if [int(x) for x in biases[:2]] != [127, 1023]:
            raise PDBError("unexpected exponent biases")
        line, _ = self._line(pos)
        chart, symtab = (int(x) for x in line.split(b"\x01")[:2])
        if not (pos + 128 <= chart <= symtab < len(b)):
            raise PDBError("chart / symbol table addresses out of range")
        self.data_start = pos + 128
        self._read_chart(chart, symtab)
        end = self._read_symtab(symtab)
        self._read_extras(end)

    def _line(self, pos):
        end = self.buf.index(b"\n", pos)
        return self.buf[pos:end], end + 1

    def _read_chart(self, pos, limit):
        self.types = {}
        while True:
            line, pos = self._line(pos)
            if line == b"\x02":
                break
            if pos > limit:
                raise PDBError("structure chart runs into the symbol table")
            f = line.split(b"\x01")
            name, size, members = f[0].decode(), int(f[1]), [m.decode() for m in f[2:] if m]
            self.types[name] = {"size": size, "members": members}
        self.chart_end = pos

    def _read_symtab(self, pos):
        self.symbols = {}
        while True:
            line, pos = self._line(pos)
            if line == b"":
                break
            f = line.split(b"\x01")
            name, typ, nitems, addr = f[0].decode(), f[1].decode(), int(f[2]), int(f[3])
            dims = [int(x) for x in f[4:] if x]
            dims = list(zip(dims[0::2], dims[1::2]))
            if dims and int(np.prod([d[1] for d in dims])) != nitems:
                raise PDBError(f"{name}: dimensions do not multiply to nitems")
            self.symbols[name] = {"type": typ, "nitems": nitems, "addr": addr, "dims": dims}
        return pos

    def _read_extras(self, pos):
        self.extras = {}
        b = self.buf
        while pos < len(b):
            line, pos = self._line(pos)
            if line == b"":
                continue
            key, _, val = line.partition(b":")
            key = key.decode()
            if key in ("Casts", "Blocks"):
                entries = []
                while True:
                    line, pos = self._line(pos)
                    if line == b"\x02":
                        break
                    entries.append(line)
                self.extras[key] = entries
            else:
                self.extras[key] = val
        al = self.extras.get("Alignment")
        if al is None or len(al) < 7:
            raise PDBError("Alignment extra missing")
        self.align = dict(zip(("char", "*", "short", "int", "long", "float", "double"), al[:7]))
        self.align["integer"] = self.align["int"]

    # ---- data -------------------------------------------------------------------------------
    _NP = {"char": "S1", "short": "<i2", "int": "<i4", "integer": "<i4", "long": "<i8", "float": "<f4", "double": "<f8"}

    def names(self):
        return list(self.symbols)

    def read(self, name):
        s = self.symbols[name]
        value, _ = self._read_items(s["type"], s["nitems"], s["addr"])
        return value

    def read_array(self, name):
        s = self.symbols[name]
        if s["type"] not in self._NP:
            raise PDBError(f"{name} is not a primitive array")
        dt = np.dtype(self._NP[s["type"]])
        if dt.itemsize != self.types[s["type"]]["size"]:
            raise PDBError("chart size disagrees with the primitive type")
        return np.frombuffer(self.buf, dtype=dt, count=s["nitems"], offset=s["addr"])

    def read_string(self, name):
        return self.read_array(name).tobytes().split(b"\0")[0].decode()

    def _layout(self, typ):
        """[(member name, base type, pointer depth, offset)], total size"""
        out, off, maxal = [], 0, 1
        for m in self.types[typ]["members"]:
            mm = re.match(r"^\s*([A-Za-z_][\w ]*?)\s*(\**)\s*([A-Za-z_]\w*)\s*$", m)
            if not mm:
                raise PDBError(f"cannot parse member '{m}'")
            base, stars, mname = mm.group(1).strip(), len(mm.group(2)), mm.group(3)
            size = self.types["*"]["size"] if stars else self.types[base]["size"]
            al = self.align["*"] if stars else self.align.get(base, 1)
            off = (off + al - 1) // al * al
            out.append((mname, base, stars, off))
            off += size
            maxal = max(maxal, al)
        total = (off + maxal - 1) // maxal * maxal
        if total != self.types[typ]["size"]:
            raise PDBError(f"struct {typ}: members add up to {total}, chart says {self.types[typ]['size']}")
        return out, total

    def _itag(self, pos):
        line, pos = self._line(pos)
        f = line.split(b"\x01")
        return int(f[0]), f[1].decode(), int(f[2]), int(f[3]), pos

    def _read_items(self, typ, nitems, pos):
        """nitems values of type `typ` stored at pos; returns (value, position after everything read)"""
        typ = typ.strip()
        if typ.endswith("*"):                      # an array of pointers: slots, then every pointee
            base = typ[:-1].strip()
            pos += nitems * self.types["*"]["size"]
            out = []
            for _ in range(nitems):
                v, pos = self._read_pointee(base, pos)
                out.append(v)
            return out, pos
        if typ in self._NP:
            dt = np.dtype(self._NP[typ])
            arr = np.frombuffer(self.buf, dtype=dt, count=nitems, offset=pos)
            pos += nitems * dt.itemsize
            if typ == "char":
                return arr.tobytes().split(b"\0")[0].decode(), pos
            return arr, pos
        layout, size = self._layout(typ)
        bodies = pos
        pos += nitems * size
        items = []
        for it in range(nitems):
            rec = {}
            for mname, base, stars, off in layout:
                if stars == 0:
                    rec[mname], _ = self._read_items(base, 1, bodies + it * size + off)
                    if not isinstance(rec[mname], str):
                        rec[mname] = rec[mname][0].item()
                else:
                    rec[mname], pos = self._read_pointee(base + " " + "*" * (stars - 1), pos)
            items.append(rec)
        return (items[0] if nitems == 1 else items), pos

    def _read_pointee(self, typ, pos):
        nitems, itype, addr, flag, pos = self._itag(pos)
        if nitems == 0 or addr == -1:
            return None, pos
        if itype.strip() != typ.strip():
            raise PDBError(f"itag says '{itype}', the structure chart says '{typ}'")
        if flag != 1:
            raise PDBError("only pointees stored in place are supported by this reader")
        return self._read_items(itype, nitems, pos)


class SiloFile(PDBFile):
    """Silo objects on top: Group -> {component: value}, literals decoded, array components read."""

    def object(self, name):
        g = self.read("/" + name)
        if g["name"] != name or g["ncomponents"] != len(g["comp_names"]) or len(g["comp_names"]) != len(g["pdb_names"]):
            raise PDBError(f"malformed Group for {name}")
        out = {"_type": g["type"]}
        for comp, val in zip(g["comp_names"], g["pdb_names"]):
            m = re.match(r"^'<([ifds])>(.*)'$", val)
            if m:
                kind, text = m.groups()
                out[comp] = int(text) if kind == "i" else float(text) if kind in "fd" else text
            else:
                s = self.symbols[val]
                out[comp] = self.read_string(val) if s["type"] == "char" else self.read_array(val)
        return out

    def objects(self):
        return [n[1:] for n, s in self.symbols.items() if s["type"] == "Group"]
