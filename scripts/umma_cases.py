"""Case generator / decoder for scripts/umma_probe.cu (development aid).

    python scripts/umma_cases.py gen  <dir>          # writes <dir>/<name>.case.bin
    python scripts/umma_cases.py decode <dir> <outdir>  # reads <outdir>/<name>.out.bin, prints the discovered maps

Discovery method: one operand is "address coded" (the bf16 at byte offset o of its region has bit pattern
0x3F80 + o/2, a distinct finite value), the other is a 0/1 selector in a layout already known to work (K-major
SWIZZLE_128B, as used by kv_proj_tc.cu).  D[r][n] then equals one coded element exactly, which tells which
shared-memory byte the tensor core read for every (row, k).
"""
import os
import struct
import sys

import numpy as np

SW_NONE, SW_128 = 0, 2


def desc(start, lbo, sbo, layout, base_offset=0):
    d = (start >> 4) & 0x3FFF
    d |= ((lbo >> 4) & 0x3FFF) << 16
    d |= ((sbo >> 4) & 0x3FFF) << 32
    d |= 1 << 46
    d |= (base_offset & 7) << 49
    d |= layout << 61
    return d


def idesc(M, N, a_major=0, b_major=0):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_major << 15) | (b_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def coded(image, off, size):
    e = np.arange(size // 2, dtype=np.uint32)
    image[off:off + size] = (0x3F80 + e).astype(np.uint16).view(np.uint8)


def put_bf16(image, off, val):
    bits = np.array([val], dtype=np.float32).view(np.uint32)[0] >> 16
    image[off:off + 2] = np.array([bits], dtype=np.uint16).view(np.uint8)


def sel_sw128(image, off, rows, fn):
    """K-major SWIZZLE_128B [rows][64]: element (r, k) = fn(r, k) for k < 16"""
    for r in range(rows):
        for k in range(16):
            v = fn(r, k)
            if v != 0:
                put_bf16(image, off + r * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2, v)


CASES = {}


def case(name):
    def deco(f):
        CASES[name] = f
        return f
    return deco


# every case returns (image, ops, ncols, info) with info = dict(coded='A'|'B', coff=region offset, M, N)
A_OFF, B_OFF = 0, 32768


def base_image():
    return np.zeros(65536, dtype=np.uint8)


def a_sel(image, M=128):
    sel_sw128(image, A_OFF, M, lambda r, k: 1.0 if (r % 16) == k else 0.0)
    return desc(A_OFF, 16, 1024, SW_128)


def b_sel(image, N=16):
    sel_sw128(image, B_OFF, N, lambda n, k: 1.0 if n == k else 0.0)
    return desc(B_OFF, 16, 1024, SW_128)


@case("c00_sanity_b_sw128")
def _():
    im = base_image()
    da = a_sel(im)
    coded(im, B_OFF, 16 * 128)
    return im, [(da, desc(B_OFF, 16, 1024, SW_128), idesc(128, 16), 0, 0)], 32, dict(coded="B", coff=B_OFF, M=128, N=16)


@case("c01_b_noswz_n16")
def _():
    im = base_image()
    da = a_sel(im)
    coded(im, B_OFF, 8192)
    return im, [(da, desc(B_OFF, 768, 1280, SW_NONE), idesc(128, 16), 0, 0)], 32, dict(coded="B", coff=B_OFF, M=128, N=16)


@case("c02_a_noswz_m128")
def _():
    im = base_image()
    db = b_sel(im)
    coded(im, A_OFF, 16384)
    return im, [(desc(A_OFF, 2064, 272, SW_NONE), db, idesc(128, 16), 0, 0)], 32, dict(coded="A", coff=A_OFF, M=128, N=16)


@case("c02b_a_noswz_m128_start48")
def _():
    im = base_image()
    db = b_sel(im)
    coded(im, A_OFF, 16384)
    return im, [(desc(A_OFF + 48, 2064, 272, SW_NONE), db, idesc(128, 16), 0, 0)], 32, dict(coded="A", coff=A_OFF, M=128, N=16)


@case("c03_a_mn_sw128_m128")
def _():
    im = base_image()
    db = b_sel(im)
    coded(im, A_OFF, 24576)
    return im, [(desc(A_OFF, 9216, 2048, SW_128), db, idesc(128, 16, a_major=1), 0, 0)], 32, dict(coded="A", coff=A_OFF, M=128, N=16)


@case("c03b_a_mn_sw128_m128_kadv")
def _():
    im = base_image()
    db = b_sel(im)
    coded(im, A_OFF, 24576)
    return im, [(desc(A_OFF + 2048, 9216, 1024, SW_128), db, idesc(128, 16, a_major=1), 0, 0)], 32, dict(coded="A", coff=A_OFF, M=128, N=16)


@case("c04_m64_lanes")
def _():
    im = base_image()
    db = b_sel(im)
    coded(im, A_OFF, 64 * 128)
    return im, [(desc(A_OFF, 16, 1024, SW_128), db, idesc(64, 16), 0, 0)], 32, dict(coded="A", coff=A_OFF, M=64, N=16)


@case("c05_m64_lane_offset16")
def _():
    im = base_image()
    db = b_sel(im)
    coded(im, A_OFF, 128 * 128)
    ops = [(desc(A_OFF, 16, 1024, SW_128), db, idesc(64, 16), 0, 0),
           (desc(A_OFF + 64 * 128, 16, 1024, SW_128), db, idesc(64, 16), (16 << 16), 0)]
    return im, ops, 32, dict(coded="A", coff=A_OFF, M=64, N=16)


@case("c06_a_mn_sw128_m64")
def _():
    im = base_image()
    db = b_sel(im)
    coded(im, A_OFF, 24576)
    return im, [(desc(A_OFF, 9216, 2048, SW_128), db, idesc(64, 16, a_major=1), 0, 0)], 32, dict(coded="A", coff=A_OFF, M=64, N=16)


@case("c07_m64_n8")
def _():
    im = base_image()
    db = b_sel(im, 8)
    coded(im, A_OFF, 64 * 128)
    return im, [(desc(A_OFF, 16, 1024, SW_128), db, idesc(64, 8), 0, 0)], 32, dict(coded="A", coff=A_OFF, M=64, N=8)


@case("c08_b_noswz_n64")
def _():
    im = base_image()
    da = a_sel(im)
    coded(im, B_OFF, 16384)
    return im, [(da, desc(B_OFF, 4112, 144, SW_NONE), idesc(128, 64), 0, 0)], 64, dict(coded="B", coff=B_OFF, M=128, N=64)


@case("c09_a_noswz_m64")
def _():
    im = base_image()
    db = b_sel(im)
    coded(im, A_OFF, 16384)
    return im, [(desc(A_OFF, 2064, 272, SW_NONE), db, idesc(64, 16), 0, 0)], 32, dict(coded="A", coff=A_OFF, M=64, N=16)


def _row_shift(shift_rows, base_off):
    im = base_image()
    db = b_sel(im)
    coded(im, A_OFF, 24576)
    start = A_OFF + 128 * shift_rows
    bo = ((start >> 7) & 7) if base_off else 0
    return im, [(desc(start, 16, 1024, SW_128, base_offset=bo), db, idesc(128, 16), 0, 0)], 32, dict(coded="A", coff=A_OFF, M=128, N=16)


@case("c10_a_sw128_rowshift1")
def _():
    return _row_shift(1, False)


@case("c10b_a_sw128_rowshift1_baseoff")
def _():
    return _row_shift(1, True)


@case("c10c_a_sw128_rowshift3")
def _():
    return _row_shift(3, False)


@case("c10d_a_sw128_rowshift3_baseoff")
def _():
    return _row_shift(3, True)


@case("c11_a_sw128_kadv_rowshift5")
def _():
    im = base_image()
    db = b_sel(im)
    coded(im, A_OFF, 24576)
    start = A_OFF + 128 * 5 + 64  # row shift 5, k step 2 (+64 B inside the 128-byte row)
    return im, [(desc(start, 16, 1024, SW_128), db, idesc(128, 16), 0, 0)], 32, dict(coded="A", coff=A_OFF, M=128, N=16)


@case("c12_b_noswz_n16_dense")
def _():
    # the layout planned for q / w operands: [k/8][n][8] with N = 16 -> LBO 256, SBO 128
    im = base_image()
    da = a_sel(im)
    coded(im, B_OFF, 8192)
    return im, [(da, desc(B_OFF, 256, 128, SW_NONE), idesc(128, 16), 0, 0)], 32, dict(coded="B", coff=B_OFF, M=128, N=16)


@case("c13_accumulate_two_ops")
def _():
    # D = A1 B^T + A2 B^T with the second op accumulating: sanity of the accumulate flag and K advance (+32 B)
    im = base_image()
    sel_sw128(im, B_OFF, 16, lambda n, k: 1.0 if n == k else 0.0)
    # second k-step of B (columns 16..31): selector again
    for n in range(16):
        k = 16 + n
        put_bf16(im, B_OFF + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2, 1.0)
    coded(im, A_OFF, 128 * 128)
    ops = [(desc(A_OFF, 16, 1024, SW_128), desc(B_OFF, 16, 1024, SW_128), idesc(128, 16), 0, 0),
           (desc(A_OFF + 32, 16, 1024, SW_128), desc(B_OFF + 32, 16, 1024, SW_128), idesc(128, 16), 0, 1)]
    return im, ops, 32, dict(coded="A2", coff=A_OFF, M=128, N=16)


def _tmem_a_case(M, N, sbo_zero=False):
    """A from tensor memory: word (lane, col) = bf16 pair coded as 0x3F80 + lane*32 + col*2 + half"""
    im = base_image()
    if N == 16 and sbo_zero:  # 8 selector rows, second 8-row group aliased onto the first (SBO = 0)
        for n in range(8):
            put_bf16(im, B_OFF + n * 16 + (n % 8) * 2, 1.0)  # no-swizzle [k/8][8 n][8]: (n, k=n) -> k chunk 0
        db = desc(B_OFF, 128, 0, SW_NONE)
    else:
        db = b_sel(im, N)
    tcols = 16
    lane = np.arange(128, dtype=np.uint32)[:, None]
    col = np.arange(tcols, dtype=np.uint32)[None, :]
    lo = 0x3F80 + lane * 32 + col * 2
    timg = (lo | ((lo + 1) << 16)).astype(np.uint32)
    ops = [(256, db, idesc(M, N), 0, 0, 0x80000000)]
    return im, ops, 32, dict(coded="TA", coff=0, M=M, N=N, timage=timg)


@case("c20_a_tmem_m128_n16")
def _():
    return _tmem_a_case(128, 16)


@case("c21_a_tmem_m64_n8")
def _():
    return _tmem_a_case(64, 8)


@case("c22_a_tmem_m128_n16_sbo0")
def _():
    return _tmem_a_case(128, 16, sbo_zero=True)


def gen(outdir):
    os.makedirs(outdir, exist_ok=True)
    for name, fn in CASES.items():
        im, ops, ncols, info = fn()
        timg = info.get("timage")
        with open(os.path.join(outdir, name + ".case.bin"), "wb") as f:
            f.write(struct.pack("<4i", im.size, len(ops), ncols, 0 if timg is None else timg.shape[1]))
            for op in ops:
                da, db, idc, dcol, acc = op[:5]
                f.write(struct.pack("<QQIIII", da, db, idc, dcol, acc, op[5] if len(op) > 5 else 0))
            f.write(im.tobytes())
            if timg is not None:
                f.write(timg.tobytes())
    print("wrote", len(CASES), "cases to", outdir)


def decode(casedir, outdir):
    for name, fn in CASES.items():
        p = os.path.join(outdir, name + ".out.bin")
        if not os.path.isfile(p):
            print(f"== {name}: no output (faulted or not run)")
            continue
        im, ops, ncols, info = fn()
        out = np.fromfile(p, dtype=np.float32).reshape(128, ncols)
        written = out != 12345.0
        lanes = np.where(written.any(axis=1))[0]
        cols = np.where(written.any(axis=0))[0]
        print(f"== {name}: lanes written {_ranges(lanes)}  cols written {_ranges(cols)}")
        bits = out.view(np.uint32)
        N = info["N"]
        if info["coded"] == "A2":
            # sum of two coded values; just print a few raw numbers
            print("   D[0][0..3] =", out[0, :4], " D[1][0..3] =", out[1, :4])
            continue
        if info["coded"] == "TA":
            code = (bits >> 16).astype(np.int64) - 0x3F80
            code[~written] = -1
            for l in [x for x in (0, 1, 15, 16, 31, 32, 33, 64, 100, 127) if x in set(lanes.tolist())]:
                print(f"   lane {l:3d}: (src lane, col, half) =", " ".join(
                    f"({int(cd) // 32},{(int(cd) % 32) // 2},{int(cd) % 2})" if cd >= 0 else "-" for cd in code[l, :N]))
            continue
        offs = ((bits >> 16).astype(np.int64) - 0x3F80) * 2
        offs[~written] = -1
        exact = (bits & 0xFFFF) == 0
        if not exact[written].all():
            print("   WARNING: some outputs are not a single coded element")
        # print the decoded byte offsets for a sample of lanes and all N columns
        show = [l for l in (0, 1, 2, 7, 8, 9, 15, 16, 17, 31, 32, 33, 47, 48, 63, 64, 65, 96, 112, 127) if l in set(lanes.tolist())]
        for l in show:
            print(f"   lane {l:3d}:", " ".join(f"{int(o):6d}" for o in offs[l, :N]))


def _ranges(a):
    a = list(map(int, a))
    if not a:
        return "[]"
    out, s, p = [], a[0], a[0]
    for x in a[1:]:
        if x != p + 1:
            out.append((s, p))
            s = x
        p = x
    out.append((s, p))
    return ",".join(f"{s}-{e}" if e > s else f"{s}" for s, e in out)


if __name__ == "__main__":
    if sys.argv[1] == "gen":
        gen(sys.argv[2])
    else:
        decode(sys.argv[2], sys.argv[3])
