"""Synthetic cohort files in the formats the reference reads (PLINK 2 .pgen/.pvar/.psam,
PLINK 1 .bed/.bim/.fam, Oxford .bgen + .bgi).

Used by the parity tests (every record type of the PGEN spec,
reference src/pgenlib_misc.h:614-905) and by bench.py (the cohort shapes of
BASELINE.json / SURVEY.md section 8d).  Writers only -- decoding is the engine's job.
"""
import os
import sqlite3
import struct
import zlib

import numpy as np

# genotype codes everywhere in this package: 0/1/2 = ALT allele count, 3 = missing
# (PLINK 2 convention, pgenlib_misc.h:619-620)


def pack_2bit(codes):
    """codes[..., N] uint8 in 0..3 -> packed [..., ceil(N/4)] uint8 (sample 0 in the low bits)."""
    codes = np.asarray(codes, np.uint8)
    n = codes.shape[-1]
    pad = (-n) % 4
    if pad:
        codes = np.concatenate([codes, np.zeros(codes.shape[:-1] + (pad,), np.uint8)], -1)
    c = codes.reshape(codes.shape[:-1] + (-1, 4))
    return (c[..., 0] | (c[..., 1] << 2) | (c[..., 2] << 4) | (c[..., 3] << 6)).astype(np.uint8)


def simulate_hardcalls(n_people, n_snps, seed, maf_range=(0.05, 0.5), missing_rate=0.0):
    """Hardy-Weinberg hardcalls, ALT frequency ~ U(maf_range). Returns codes[M, N] uint8."""
    rng = np.random.default_rng(seed)
    p = rng.uniform(maf_range[0], maf_range[1], size=(n_snps, 1))
    u = rng.random((n_snps, n_people))
    codes = np.where(u < p * p, 2, np.where(u < p * p + 2 * p * (1 - p), 1, 0)).astype(np.uint8)
    if missing_rate > 0:
        codes[rng.random((n_snps, n_people)) < missing_rate] = 3
    return codes


# ----------------------------------------------------------------------------- PGEN

def _varint(v):
    out = bytearray()
    while v > 127:
        out.append((v & 127) | 128)
        v >>= 7
    out.append(v)
    return bytes(out)


def _id_bytes(n):
    b = 1
    while b < 4 and (n >> (8 * b)):
        b += 1
    return b


def _sample_list(ids, n_samples, values=None):
    """Difflist (values given) or deltalist (values None), pgenlib_misc.h:699-718."""
    ids = [int(i) for i in ids]
    out = bytearray(_varint(len(ids)))
    if not ids:
        return bytes(out)
    assert len(ids) <= n_samples // 8, "difflist longer than N/8"
    ib = _id_bytes(n_samples)
    groups = [ids[k:k + 64] for k in range(0, len(ids), 64)]
    deltas = [b"".join(_varint(g[j] - g[j - 1]) for j in range(1, len(g))) for g in groups]
    for g in groups:
        out += int(g[0]).to_bytes(ib, "little")
    for d in deltas[:-1]:
        out.append(len(d) - 63)
    if values is not None:
        out += pack_2bit(np.asarray(values, np.uint8)).tobytes()
    for d in deltas:
        out += d
    return bytes(out)


def _invert(codes):
    c = np.asarray(codes, np.uint8).copy()
    c[codes == 0] = 2
    c[codes == 2] = 0
    return c


def encode_main_track(codes, kind, n_samples, base=None):
    """Encode one variant's hardcalls with main-track type `kind`
    (0 raw, 1 one-bit, 2 LD, 3 inverted LD, 4/6/7 difflist on constant, 5 all-ref)."""
    codes = np.asarray(codes, np.uint8)
    if kind == 0:
        return pack_2bit(codes).tobytes()
    if kind == 5:
        assert not codes.any()
        return b""
    if kind in (4, 6, 7):
        ids = np.nonzero(codes != (kind & 3))[0]
        return _sample_list(ids, n_samples, codes[ids])
    if kind == 1:
        counts = np.bincount(codes, minlength=4)
        lo, hi = sorted(np.argsort(-counts, kind="stable")[:2])
        bits = np.packbits((codes == hi).astype(np.uint8), bitorder="little").tobytes()
        ids = np.nonzero((codes != lo) & (codes != hi))[0]
        return bytes([(int(lo) << 2) | int(hi - lo)]) + bits + _sample_list(ids, n_samples, codes[ids])
    if kind in (2, 3):
        target = codes if kind == 2 else _invert(codes)
        ids = np.nonzero(target != base)[0]
        return _sample_list(ids, n_samples, target[ids])
    raise ValueError(kind)


def encode_dosage_track(dosage, kind, n_samples):
    """dosage[N] uint16 (0xFFFF = none). kind 0x20 list, 0x40 unconditional, 0x60 bitarray."""
    dosage = np.asarray(dosage, np.uint16)
    have = dosage != 0xFFFF
    if kind == 0x40:
        return dosage.astype("<u2").tobytes()
    if kind == 0x60:
        return np.packbits(have.astype(np.uint8), bitorder="little").tobytes() + dosage[have].astype("<u2").tobytes()
    if kind == 0x20:
        return _sample_list(np.nonzero(have)[0], n_samples) + dosage[have].astype("<u2").tobytes()
    raise ValueError(kind)


def encode_hphase_track(codes, rng, explicit):
    het = int((np.asarray(codes) == 1).sum())
    if not explicit:
        bits = np.concatenate([[0], rng.integers(0, 2, het)]).astype(np.uint8)
        return np.packbits(bits, bitorder="little").tobytes().ljust(1 + het // 8, b"\0")
    present = rng.integers(0, 2, het).astype(np.uint8)
    present[rng.integers(0, het)] = 1  # an explicit track with no phased het is malformed
    first = np.packbits(np.concatenate([[1], present]).astype(np.uint8), bitorder="little").tobytes()
    first = first.ljust(1 + het // 8, b"\0")
    info = np.packbits(rng.integers(0, 2, int(present.sum())).astype(np.uint8), bitorder="little").tobytes()
    return first + info


def write_pgen_fixed(path, codes):
    """Mode 0x02: header + M records of ceil(N/4) bytes (pgenlib_misc.h:619-620).  Header control
    byte 0: nonref flags unstored -- the reference provider passes nonref_flags_already_loaded=1 with
    no flag array (src/loader.cpp:323), so fixed-width files that claim "all ref" / "never ref"
    make pgenlib dereference a null pointer (pgenlib_read.cc:1081-1101)."""
    codes = np.asarray(codes, np.uint8)
    m, n = codes.shape
    with open(path, "wb") as f:
        f.write(b"\x6c\x1b\x02" + struct.pack("<II", m, n) + b"\x00")
        f.write(pack_2bit(codes).tobytes())


def write_pgen_fixed_dosage(path, codes, dosage):
    """Mode 0x03: 2-bit hardcalls followed by N u16 dosages per variant."""
    codes = np.asarray(codes, np.uint8)
    m, n = codes.shape
    with open(path, "wb") as f:
        f.write(b"\x6c\x1b\x03" + struct.pack("<II", m, n) + b"\x00")
        packed = pack_2bit(codes)
        for i in range(m):
            f.write(packed[i].tobytes() + np.asarray(dosage[i], "<u2").tobytes())


def write_pgen_variable(path, codes, kinds, dosage=None, dosage_kinds=None, hphase=None, seed=0,
                        len_bytes=None):
    """Mode 0x10 file.  kinds[i] = main-track type for variant i (LD types take the most
    recent non-LD variant as base); dosage_kinds[i] in {0,0x20,0x40,0x60}; hphase[i] in
    {0, 1 (implicit), 2 (explicit phasepresent)}."""
    codes = np.asarray(codes, np.uint8)
    m, n = codes.shape
    rng = np.random.default_rng(seed)
    records, vrtypes = [], []
    base = None
    for i in range(m):
        k = int(kinds[i])
        if k in (2, 3):
            assert base is not None and i % 65536 != 0
            rec = encode_main_track(codes[i], k, n, base)
        else:
            rec = encode_main_track(codes[i], k, n)
            base = codes[i]
        vt = k
        if hphase is not None and hphase[i] and (codes[i] == 1).any():
            rec += encode_hphase_track(codes[i], rng, hphase[i] == 2)
            vt |= 0x10
        if dosage_kinds is not None and dosage_kinds[i]:
            rec += encode_dosage_track(dosage[i], int(dosage_kinds[i]), n)
            vt |= int(dosage_kinds[i])
        records.append(rec)
        vrtypes.append(vt)
    wide = any(v > 15 for v in vrtypes)
    maxlen = max(len(r) for r in records)
    if len_bytes is None:
        len_bytes = 1 if maxlen < 256 else 2 if maxlen < 65536 else 3 if maxlen < (1 << 24) else 4
    ctrl = (len_bytes - 1) | (4 if wide else 0) | 0x80
    vblocks = (m + 65535) // 65536
    hdr = bytearray(b"\x6c\x1b\x10" + struct.pack("<II", m, n) + bytes([ctrl]))
    blocks = bytearray()
    for b in range(vblocks):
        vt = vrtypes[b * 65536:(b + 1) * 65536]
        rl = [len(r) for r in records[b * 65536:(b + 1) * 65536]]
        if wide:
            blocks += bytes(vt)
        else:
            vt2 = vt + [0] * (len(vt) & 1)
            blocks += bytes(vt2[j] | (vt2[j + 1] << 4) for j in range(0, len(vt2), 2))
        for x in rl:
            blocks += int(x).to_bytes(len_bytes, "little")
    first = len(hdr) + 8 * vblocks + len(blocks)
    pos = first
    for b in range(vblocks):
        hdr += struct.pack("<Q", pos)
        pos += sum(len(r) for r in records[b * 65536:(b + 1) * 65536])
    with open(path, "wb") as f:
        f.write(bytes(hdr) + bytes(blocks))
        for r in records:
            f.write(r)
    return vrtypes


def write_bed(path, codes):
    """PLINK 1 variant-major .bed: 00 hom A1(=2 copies), 01 missing, 10 het, 11 hom A2."""
    codes = np.asarray(codes, np.uint8)
    bed = np.array([3, 2, 0, 1], np.uint8)[codes]
    with open(path, "wb") as f:
        f.write(b"\x6c\x1b\x01")
        f.write(pack_2bit(bed).tobytes())


def write_pvar(path, n_snps, chrom="1"):
    with open(path, "w") as f:
        f.write("#CHROM\tPOS\tID\tREF\tALT\n")
        for i in range(n_snps):
            f.write(f"{chrom}\t{1000 * (i + 1)}\tRSID_{i + 1}\tG\tA\n")


def write_psam(path, n_people):
    with open(path, "w") as f:
        f.write("#FID\tIID\tSEX\n")
        for i in range(n_people):
            f.write(f"fam{i + 1}\tid{i + 1}\tNA\n")


def write_bim(path, n_snps, chrom="1"):
    with open(path, "w") as f:
        for i in range(n_snps):
            f.write(f"{chrom}\tRSID_{i + 1}\t0\t{1000 * (i + 1)}\tA\tG\n")


def write_fam(path, n_people):
    with open(path, "w") as f:
        for i in range(n_people):
            f.write(f"fam{i + 1} id{i + 1} 0 0 0 -9\n")


# ----------------------------------------------------------------------------- BGEN

def _lstr(s, width):
    b = s.encode()
    return len(b).to_bytes(width, "little") + b


def write_bgen(path, prob_ints, layout=2, bits=8, compression=1, positions=None, chrom="01",
               missing=None, phased=False, write_index=True):
    """prob_ints[M, N, 2] integers: numerators of P(first-allele hom), P(het).
    layout 1: u16 numerators over 32768 (a third one is derived); layout 2: `bits`-bit numerators
    over 2^bits-1.  missing[M, N] bool sets the layout-2 ploidy-byte missing flag.  Writes
    `<path>.bgi` (the sqlite index the reference requires, src/bgen.cpp:506-664)."""
    prob_ints = np.asarray(prob_ints, np.uint64)
    m, n, _ = prob_ints.shape
    if positions is None:
        positions = [1000 * (i + 1) for i in range(m)]
    flags = (compression & 3) | (layout << 2)
    header = struct.pack("<III", 20, m, n) + b"bgen" + struct.pack("<I", flags)
    index_rows = []
    body = bytearray()
    start0 = 4 + len(header)
    for i in range(m):
        start = start0 + len(body)
        rec = bytearray()
        if layout == 1:
            rec += struct.pack("<I", n)
        rec += _lstr(f"SNPID_{i + 1}", 2) + _lstr(f"RSID_{i + 1}", 2) + _lstr(chrom, 2)
        rec += struct.pack("<I", positions[i])
        if layout == 2:
            rec += struct.pack("<H", 2)
        rec += _lstr("A", 4) + _lstr("G", 4)
        if layout == 1:
            p = prob_ints[i]
            third = 32768 - p[:, 0] - p[:, 1]
            raw = np.stack([p[:, 0], p[:, 1], third], 1).astype("<u2").tobytes()
            if compression:
                z = zlib.compress(raw)
                rec += struct.pack("<I", len(z)) + z
            else:
                rec += raw
        else:
            pl = np.full(n, 2, np.uint8)
            if missing is not None:
                pl = pl | (np.asarray(missing[i], np.uint8) << 7)
            vals = prob_ints[i].reshape(-1)
            if bits == 8:
                packed = vals.astype(np.uint8).tobytes()
            elif bits == 16:
                packed = vals.astype("<u2").tobytes()
            else:
                acc, nb, out = 0, 0, bytearray()
                for v in vals.tolist():
                    acc |= int(v) << nb
                    nb += bits
                    while nb >= 8:
                        out.append(acc & 255)
                        acc >>= 8
                        nb -= 8
                if nb:
                    out.append(acc & 255)
                packed = bytes(out)
            raw = struct.pack("<IHBB", n, 2, 2, 2) + pl.tobytes() + bytes([1 if phased else 0, bits]) + packed
            if compression:
                z = zlib.compress(raw)
                rec += struct.pack("<II", len(z) + 4, len(raw)) + z
            else:
                rec += struct.pack("<I", len(raw)) + raw
        body += rec
        index_rows.append((chrom, positions[i], f"RSID_{i + 1}", 2, "A", "G", start, len(rec)))
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(header)) + header + bytes(body))
    if write_index:
        bgi = path + ".bgi"
        if os.path.exists(bgi):
            os.remove(bgi)
        con = sqlite3.connect(bgi)
        con.execute("CREATE TABLE Variant (chromosome TEXT NOT NULL, position INT NOT NULL, rsid TEXT NOT NULL, "
                    "number_of_alleles INT NOT NULL, allele1 TEXT NOT NULL, allele2 TEXT NULL, "
                    "file_start_position INT NOT NULL, size_in_bytes INT NOT NULL, "
                    "PRIMARY KEY (chromosome, position, rsid, allele1, allele2, file_start_position)) WITHOUT ROWID")
        con.executemany("INSERT INTO Variant VALUES (?,?,?,?,?,?,?,?)", index_rows)
        con.commit()
        con.close()


# ------------------------------------------------------------------- device-side cohorts

def device_hardcalls(n_people, n_snps, seed, device="cuda", maf_range=(0.05, 0.5), missing_rate=0.0,
                     chunk=2048):
    """Hardy-Weinberg hardcalls generated on the GPU, packed 2 bits per person in the engine's
    HBM layout: uint8 tensor [n_snps, pitch], pitch = ceil(N/4) rounded up to 16 bytes, PLINK 2
    coding.  (torch is used as the RNG / allocator only.)"""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    row_bytes = (n_people + 3) // 4
    pitch = (row_bytes + 15) // 16 * 16
    out = torch.zeros((n_snps, pitch), dtype=torch.uint8, device=device)
    n4 = row_bytes * 4
    for s0 in range(0, n_snps, chunk):
        m = min(chunk, n_snps - s0)
        p = torch.empty((m, 1), device=device).uniform_(maf_range[0], maf_range[1], generator=g)
        u = torch.rand((m, n4), device=device, generator=g)
        codes = (u < p * p).to(torch.uint8) * 2 + ((u >= p * p) & (u < p * p + 2 * p * (1 - p))).to(torch.uint8)
        if missing_rate > 0:
            miss = torch.rand((m, n4), device=device, generator=g) < missing_rate
            codes[miss] = 3
        codes[:, n_people:] = 0
        c = codes.view(m, row_bytes, 4)
        out[s0:s0 + m, :row_bytes] = c[..., 0] | (c[..., 1] << 2) | (c[..., 2] << 4) | (c[..., 3] << 6)
    return out
