"""CPU study (NumPy emulation of the operand roundings, fp32 accumulation) of cheaper hi / lo split schemes for the actor:
which of them keep the probabilities as close to an fp64 forward as the shipped 3xTF32 kernel does?

    python profiles/tools/actor_split_study.py      # no GPU needed

Result (profiles/r2_actor_split_study.txt): three passes on fp16 halves (kind::f16, K = 16 per instruction: HALF the
instruction count of 3 x tf32) with the lo halves scaled by 2^11 and the matching hi halves scaled by 2^-11 -- so that all
three passes add into ONE accumulator -- are as accurate as 3 x tf32, also when the scaled-down halves go subnormal."""
import numpy as np, torch
torch.manual_seed(3)
def tf32(x):  # round to nearest, ties away (bits + 0x1000) & ~0x1fff
    b = x.astype(np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xffffe000)).view(np.float32)
def bf16(x):
    b = x.astype(np.float32).view(np.uint32)
    return ((b + np.uint32(0x8000)) & np.uint32(0xffff0000)).view(np.float32)
def fp16(x):
    return x.astype(np.float16).astype(np.float32)
def mm32(a, b):   # fp32 accumulate (numpy float32 matmul accumulates in fp32 pairwise-ish; fine for a study)
    return (a.astype(np.float32) @ b.astype(np.float32)).astype(np.float32)
D, H = 50, 100
fc = [torch.nn.Linear(D, H), torch.nn.Linear(H, H), torch.nn.Linear(H, 2)]
with torch.no_grad():
    for l in fc: l.weight.mul_(3.0)
W = [l.weight.detach().numpy().T.astype(np.float32) for l in fc]; B = [l.bias.detach().numpy().astype(np.float32) for l in fc]
rng = np.random.default_rng(0)
X = rng.normal(0, 1, (20000, D)).astype(np.float32) * np.array([1, 1, 4, 1, 3, 3, .5, 1, 1, 1] * 5, dtype=np.float32)
def forward(split):
    a = X
    for li in range(2):
        z = split(a, W[li]) + B[li]
        a = np.maximum(z, 0).astype(np.float32)
    z = a.astype(np.float32) @ W[2] + B[2]    # output layer in fp32 FMAs
    e = np.exp(z - z.max(1, keepdims=True)); return (e / e.sum(1, keepdims=True))[:, 1]
def ref64():
    a = X.astype(np.float64)
    for li in range(2): a = np.maximum(a @ W[li].astype(np.float64) + B[li], 0)
    z = a @ W[2].astype(np.float64) + B[2]
    e = np.exp(z - z.max(1, keepdims=True)); return (e / e.sum(1, keepdims=True))[:, 1]
def s_fp32(a, w): return mm32(a, w)
def s_tf32(a, w): return mm32(tf32(a), tf32(w))
def s_3x(a, w):
    ah, wh = tf32(a), tf32(w); al, wl = tf32(a - ah), tf32(w - wh)
    return mm32(ah, wh) + mm32(al, wh) + mm32(ah, wl)
def s_tf32_bf16corr(a, w):   # main pass tf32, the two correction passes in bf16 (K = 16 per instruction)
    ah, wh = tf32(a), tf32(w); al, wl = a - ah, w - wh
    return mm32(ah, wh) + mm32(bf16(al), bf16(wh)) + mm32(bf16(ah), bf16(wl))
def s_tf32_fp16corr(a, w):   # corrections in fp16 with the lo halves scaled by 2^11 (range), result scaled back
    ah, wh = tf32(a), tf32(w); al, wl = a - ah, w - wh
    sc = np.float32(2048.0)
    return mm32(ah, wh) + (mm32(fp16(al * sc), fp16(wh)) + mm32(fp16(ah), fp16(wl * sc))) / sc
def s_fp16x3(a, w):          # everything in fp16 halves (K = 16 for all three passes), lo halves scaled
    sc = np.float32(2048.0)
    ah, wh = fp16(a), fp16(w); al, wl = fp16((a - ah) * sc), fp16((w - wh) * sc)
    return mm32(ah, wh) + (mm32(al, wh) + mm32(ah, wl)) / sc
r = ref64()
for name, f in (("fp32 forward", s_fp32), ("1 x tf32", s_tf32), ("3 x tf32 (shipped)", s_3x), ("tf32 + 2 x bf16 corrections", s_tf32_bf16corr),
                ("tf32 + 2 x fp16 corrections (lo scaled 2^11)", s_tf32_fp16corr), ("3 x fp16 (lo scaled 2^11)", s_fp16x3)):
    p = forward(f); print(f"{name:48s} max |dp| = {np.abs(p - r).max():.2e}   mean = {np.abs(p - r).mean():.2e}")
def s_fp16x3_one_acc(a, w):  # ONE accumulator: the 2^-11 rides on an operand of each correction pass
    sc = np.float32(2048.0)
    ah, wh = fp16(a), fp16(w)
    al_s, wl_s = fp16((a - ah) * sc), fp16((w - wh) * sc)          # lo halves scaled up
    ah_d, wh_d = fp16(ah / sc), fp16(wh / sc)                        # hi halves scaled down (may go subnormal)
    return mm32(ah, wh) + mm32(al_s, wh_d) + mm32(ah_d, wl_s)
p = forward(s_fp16x3_one_acc); print(f"{'3 x fp16, one accumulator (operands pre-scaled)':48s} max |dp| = {np.abs(p - r).max():.2e}   mean = {np.abs(p - r).mean():.2e}")
# weights ten times smaller and observations ten times larger: the scaled-down hi halves go subnormal
W = [w * np.float32(0.1) for w in W]; X = X * np.float32(10.0); r = ref64()
for name, f in (("[small weights] 3 x tf32", s_3x), ("[small weights] 3 x fp16 one accumulator", s_fp16x3_one_acc), ("[small weights] fp32 forward", s_fp32)):
    p = forward(f); print(f"{name:48s} max |dp| = {np.abs(p - r).max():.2e}   mean = {np.abs(p - r).mean():.2e}")
