"""Which fp32 operation order does torch CPU grid_sample (bilinear, border, align_corners=False) execute? Emulates the
candidates in numpy (FMA via long double) and counts bit mismatches against torch itself. Result on torch 2.11 (AVX2 /
AVX-512 build): unnormalise = fma(g + 1, size / 2, -0.5), sum = fma(se, w_se, fma(sw, w_sw, fma(ne, w_ne, nw * w_nw))) ->
0 / 60000 mismatches; pi3b_keypoint_gather restates exactly that."""
import torch, numpy as np
torch.manual_seed(0)
H, W, K = 308, 406, 20000
img = torch.rand(1, 3, H, W)
kx = torch.rand(K) * (W - 1); ky = torch.rand(K) * (H - 1)
# include integer-ish keypoints like the grid extractor
kx[:5000] = torch.randint(0, W, (5000,)).float(); ky[:5000] = torch.randint(0, H, (5000,)).float()
gx = (kx / (W - 1)) * 2 - 1; gy = (ky / (H - 1)) * 2 - 1
grid = torch.stack([gx, gy], -1)[None, None]
ref = torch.nn.functional.grid_sample(img, grid, mode="bilinear", align_corners=False, padding_mode="border")[0, :, 0].T.contiguous()  # K,3
f32 = np.float32
def fma(a, b, c):
    return (a.astype(np.longdouble) * b.astype(np.longdouble) + c.astype(np.longdouble)).astype(np.float32)
gxn, gyn = gx.numpy(), gy.numpy()
im = img[0].numpy()
def run(unnorm, wmode, summode):
    def un(g, size):
        if unnorm == "scalar":   # ((g+1)*size-1)/2
            return ((g + f32(1)) * f32(size) - f32(1)) / f32(2)
        if unnorm == "vec":      # (g+1)*(size/2) - 0.5
            return (g + f32(1)) * f32(size / 2) - f32(0.5)
        if unnorm == "vecfma":
            return fma(g + f32(1), np.full_like(g, f32(size / 2)), np.full_like(g, f32(-0.5)))
    x = np.clip(un(gxn, W), f32(0), f32(W - 1)); y = np.clip(un(gyn, H), f32(0), f32(H - 1))
    xw = np.floor(x); yn = np.floor(y)
    if wmode == "vec":
        w = x - xw; e = f32(1) - w; n = y - yn; s = f32(1) - n
    else:
        w = x - xw; e = (xw + f32(1)) - x; n = y - yn; s = (yn + f32(1)) - y
    nw, ne, sw, se = s * e, s * w, n * e, n * w
    x0 = xw.astype(np.int64); y0 = yn.astype(np.int64); x1 = x0 + 1; y1 = y0 + 1
    def val(yy, xx):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        v = im[:, np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
        return np.where(ok, v, f32(0)).astype(np.float32)
    a, b, c, d = val(y0, x0), val(y0, x1), val(y1, x0), val(y1, x1)
    if summode == "plain":
        out = ((a * nw + b * ne) + c * sw) + d * se
    elif summode == "fma_l":   # fma(d,se, fma(c,sw, fma(b,ne, a*nw)))
        out = fma(d, np.broadcast_to(se, d.shape), fma(c, np.broadcast_to(sw, c.shape), fma(b, np.broadcast_to(ne, b.shape), a * nw)))
    elif summode == "fma_r":   # fma(a,nw, b*ne) then fma(c,sw,.) fma(d,se,.)
        out = fma(d, np.broadcast_to(se, d.shape), fma(c, np.broadcast_to(sw, c.shape), fma(a, np.broadcast_to(nw, a.shape), b * ne)))
    return out.T
r = ref.numpy()
for unnorm in ("scalar", "vec", "vecfma"):
    for wmode in ("vec", "scalar"):
        for summode in ("plain", "fma_l", "fma_r"):
            o = run(unnorm, wmode, summode)
            nbad = int((o != r).sum())
            u8bad = int(((o * f32(255)).astype(np.uint8) != (r * f32(255)).astype(np.uint8)).sum())
            print(f"{unnorm:7s} {wmode:6s} {summode:6s} float mismatches {nbad:6d} / {r.size}   uint8 mismatches {u8bad}")
