"""Circom `.r1cs` reader (host-side data format of row f4).

Mirrors `legogroth16::circom::R1CSFile::new` / `Header::new` (legogroth16/src/circom/r1cs_reader.rs:33-230,
structures in legogroth16/src/circom/r1cs.rs:9-95) for circuits compiled with `-p bls12381`: same field
names, same acceptance rules and the same `R1CSFileParsing` error messages.  Binary format: iden3 r1csfile
(magic "r1cs", version 1, typed sections: 1 header, 2 constraints, 3 wire-to-label map; sections may come in
any order; all integers little-endian).

On top of the reader, `R1CSFile.matrices()` lays the three constraint matrices out as CSR arrays with
canonical 32-byte little-endian coefficients - the layout a device-side sparse A.w / B.w / C.w product
consumes before `dg_qap_h_from_abc` (legogroth16/src/r1cs_to_qap.rs:150-207) - and `evaluate()` is the
host-side definition of that product (Python integers; small circuits / tests only).
"""
import io
import struct

import numpy as np

from .msm import R_MODULUS

MAGIC = b'r1cs'


class CircomError(Exception):
    """legogroth16::circom::error::CircomError::{UnableToOpenR1CSFile, R1CSFileParsing}."""

    def __init__(self, kind, msg):
        super().__init__('%s: %s' % (kind, msg))
        self.kind, self.msg = kind, msg


def _parsing(msg):
    return CircomError('R1CSFileParsing', msg)


def _read_exact(r, n):
    b = r.read(n)
    if len(b) != n:
        raise _parsing('Encountered error while parsing R1CS file: failed to fill whole buffer')
    return b


def _u32(r):
    return struct.unpack('<I', _read_exact(r, 4))[0]


def _u64(r):
    return struct.unpack('<Q', _read_exact(r, 8))[0]


class Header:
    """field_size, subgroup_order (LE bytes), curve, n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_constraints."""

    def __init__(self, reader, size):
        self.field_size = _u32(reader)
        if self.field_size != 32:
            raise _parsing('This parser only supports 32-byte fields')
        if size != 32 + self.field_size:
            raise _parsing('Invalid header section size')
        self.subgroup_order = _read_exact(reader, self.field_size)
        if int.from_bytes(self.subgroup_order, 'little') != R_MODULUS:      # check_subgroup_order: this backend is BLS12-381 only
            raise CircomError('IncompatibleWithCurve', 'the circuit was not compiled for BLS12-381 (-p bls12381)')
        self.curve = 'Bls12_381'
        self.n_wires = _u32(reader)
        self.n_pub_out = _u32(reader)
        self.n_pub_in = _u32(reader)
        self.n_prv_in = _u32(reader)
        self.n_labels = _u64(reader)
        self.n_constraints = _u32(reader)


def _read_lc(reader):
    terms = []
    for _ in range(_u32(reader)):
        wire = _u32(reader)
        coeff = int.from_bytes(_read_exact(reader, 32), 'little')
        if coeff >= R_MODULUS:                      # Fr::deserialize_uncompressed rejects non-canonical values
            raise _parsing('Encountered error while parsing R1CS file: InvalidData')
        terms.append((wire, coeff))
    return terms


class R1CSFile:
    """version, header, constraints (list of (a, b, c), each a list of (wire_id, coefficient)), wire_mapping."""

    def __init__(self, version, header, constraints, wire_mapping):
        self.version, self.header, self.constraints, self.wire_mapping = version, header, constraints, wire_mapping

    @classmethod
    def new_from_file(cls, path):
        try:
            f = open(path, 'rb')
        except OSError as e:
            raise CircomError('UnableToOpenR1CSFile', 'Encountered error while opening R1CS file: %r' % (e,))
        with f:
            return cls.new(f)

    @classmethod
    def new(cls, reader):
        if isinstance(reader, (bytes, bytearray, memoryview)):
            reader = io.BytesIO(bytes(reader))
        if _read_exact(reader, 4) != MAGIC:
            raise _parsing('Invalid magic number')
        version = _u32(reader)
        if version != 1:
            raise _parsing('Unsupported version')
        offsets, sizes = {}, {}
        for _ in range(_u32(reader)):
            t, size = _u32(reader), _u64(reader)
            offsets[t], sizes[t] = reader.tell(), size
            try:
                reader.seek(size, io.SEEK_CUR)
            except (OverflowError, OSError, ValueError):
                raise _parsing('Invalid section size') from None
        for t, name in ((1, 'header'), (2, 'constraint'), (3, 'wire2label')):
            if t not in offsets:
                raise _parsing('No section offset for %s type found' % name)
        reader.seek(offsets[1])
        header = Header(reader, sizes[1])
        reader.seek(offsets[2])
        constraints = [(_read_lc(reader), _read_lc(reader), _read_lc(reader)) for _ in range(header.n_constraints)]
        reader.seek(offsets[3])
        if sizes[3] != header.n_wires * 8:
            raise _parsing('Invalid map section size')
        wire_mapping = [_u64(reader) for _ in range(header.n_wires)]
        if wire_mapping and wire_mapping[0] != 0:                    # read_map, r1cs_reader.rs:233
            raise _parsing('Wire 0 should always be mapped to 0')
        return cls(version, header, constraints, wire_mapping)

    # ---- derived views ------------------------------------------------------------------------------------
    @property
    def num_public(self):
        """R1CS::num_public: public outputs and inputs plus the constant wire 1 (legogroth16/src/circom/r1cs.rs:47-49)."""
        return 1 + self.header.n_pub_out + self.header.n_pub_in

    @property
    def num_private(self):
        return self.header.n_wires - self.num_public

    def matrices(self):
        """[(row_ptr u32[m+1], col u32[nnz], coeff u8[nnz x 32])] for A, B, C; coefficients canonical little-endian."""
        out = []
        for k in range(3):
            row_ptr, cols, vals = [0], [], []
            for con in self.constraints:
                for wire, coeff in con[k]:
                    if wire >= self.header.n_wires:
                        raise _parsing('wire index out of range')
                    cols.append(wire)
                    vals.append(coeff.to_bytes(32, 'little'))
                row_ptr.append(len(cols))
            out.append((np.asarray(row_ptr, np.uint32), np.asarray(cols, np.uint32),
                        np.frombuffer(b''.join(vals), np.uint8).reshape(-1, 32) if vals else np.zeros((0, 32), np.uint8)))
        return out

    def evaluate(self, witness):
        """(A.w, B.w, C.w) mod r as lists of ints; witness[0] must be 1.  Host-side definition of the sparse products."""
        if len(witness) != self.header.n_wires:
            raise ValueError('witness has %d wires, the circuit %d' % (len(witness), self.header.n_wires))
        res = ([], [], [])
        for con in self.constraints:
            for k in range(3):
                res[k].append(sum(c * witness[w] for w, c in con[k]) % R_MODULUS)
        return res

    def is_satisfied(self, witness):
        a, b, c = self.evaluate(witness)
        return all((x * y - z) % R_MODULUS == 0 for x, y, z in zip(a, b, c))


def write_r1cs(n_pub_out, n_pub_in, n_prv_in, n_wires, constraints, wire_mapping=None, n_labels=None, prime=R_MODULUS,
               section_order=(1, 2, 3)):
    """Serialises a circuit in the iden3 format (test fixtures; the inverse of R1CSFile.new)."""
    wire_mapping = list(range(n_wires)) if wire_mapping is None else wire_mapping
    n_labels = n_wires if n_labels is None else n_labels
    header = (struct.pack('<I', 32) + int(prime).to_bytes(32, 'little') +
              struct.pack('<IIIIQI', n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, len(constraints)))

    def lc(terms):
        return struct.pack('<I', len(terms)) + b''.join(struct.pack('<I', w) + int(c % prime).to_bytes(32, 'little') for w, c in terms)
    body = {1: header, 2: b''.join(lc(a) + lc(b) + lc(c) for a, b, c in constraints),
            3: b''.join(struct.pack('<Q', x) for x in wire_mapping)}
    out = MAGIC + struct.pack('<II', 1, len(section_order))
    for t in section_order:
        out += struct.pack('<IQ', t, len(body[t])) + body[t]
    return out
