"""CPU model of the index arithmetic of cost_volume_split_row32_kernel (cost_volume_tcs.cu): the band GEMM per 4 x 32
tile is formed with numpy, then the epilogue's addressing (TMEM column of a candidate row, 40-word row buffer, pick at
lane + dh, 81-float output rows, the 21-unit float4 output loop) is replayed literally and compared with the oracle's
cost volume.  Checks the logic that could be checked without a GPU; barriers / descriptors are not modelled."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import pwc_oracle as O

R_TW, R_TH, R_FW, R_FH = 32, 4, 40, 12
R_NH = R_FW * (R_FH // 2)
IN_PITCH, OUT_PITCH = 44, 84


def patch(f, b, y0, x0, h, w):
    """TMA box with zero fill outside the image: rows y0..y0+h-1, cols x0..x0+w-1 -> (h*w, C), x fastest."""
    B, H, W, C = f.shape
    out = np.zeros((h, w, C), np.float64)
    for r in range(h):
        for c in range(w):
            y, x = y0 + r, x0 + c
            if 0 <= y < H and 0 <= x < W:
                out[r, c] = f[b, y, x]
    return out.reshape(h * w, C)


def run(shape, alpha=0.1, out_cs=88, seed=0):
    B, H, W, C = shape
    rng = np.random.default_rng(seed)
    f0 = rng.standard_normal(shape).astype(np.float32)
    f1 = rng.standard_normal(shape).astype(np.float32)
    ref = O.cost_volume(torch.from_numpy(f0), torch.from_numpy(f1), 4).numpy()
    out = np.full((B, H, W, out_cs), 7.0, np.float32)
    scale = 1.0 / C
    tiles_x, tiles_y = -(-W // R_TW), -(-H // R_TH)
    for t in range(tiles_x * tiles_y * B):
        tx, ty, b = t % tiles_x, (t // tiles_x) % tiles_y, t // (tiles_x * tiles_y)
        x0, y0 = tx * R_TW, ty * R_TH
        A = patch(f0, b, y0, x0, R_TH, R_TW)                   # 128 rows: m = py*32 + px
        Bm = patch(f1, b, y0 - 4, x0 - 4, R_FH, R_FW)           # 480 rows: n = fy*40 + fx
        D = A @ Bm.T                                            # TMEM: lane m, column n (half hb at column hb*240)
        for q in range(4):                                      # epilogue warp = quadrant = tile row
            yy = y0 + q
            in_slab = np.zeros((32, IN_PITCH)); out_slab = np.zeros((32, OUT_PITCH))
            for dv in range(9):
                r = q + dv; hb = 1 if r >= R_FH // 2 else 0
                ta = hb * R_NH + (r - hb * (R_FH // 2)) * R_FW   # column of the candidate row inside the accumulator
                for lane in range(32):
                    in_slab[lane, :40] = D[q * 32 + lane, ta:ta + 40]
                for lane in range(32):
                    for dh in range(9):
                        out_slab[lane, dv * 9 + dh] = in_slab[lane, lane + dh]
            if yy < H:
                npx = min(32, W - x0)
                for lane in range(32):                          # the float4 output loop, unit u = lane + 32 m
                    pix, k = lane // 21, lane % 21
                    for m in range(21):
                        if pix < npx:
                            n = 4 if k < 20 else 1
                            v = out_slab[pix, 4 * k:4 * k + n] * scale
                            out[b, yy, x0 + pix, 4 * k:4 * k + n] = np.maximum(v, alpha * v)
                        k += 11; pix += 1
                        if k >= 21:
                            k -= 21; pix += 1
    err = float(np.abs(out[..., :81] - ref).max())
    untouched = float(np.abs(out[..., 81:] - 7.0).max())
    print(shape, "max|err|", err, "slot overrun", untouched)
    assert err < 1e-5 and untouched == 0


if __name__ == "__main__":
    for s in [(1, 8, 64, 32), (2, 13, 70, 32), (1, 5, 31, 64)]:
        run(s)
    print("row32 index model: ok")
