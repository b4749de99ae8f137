#!/usr/bin/env python3
"""profiles/r02_k3_sass_excerpt.md from the built library: python profiles/make_k3_sass_excerpt.py > profiles/r02_k3_sass_excerpt.md"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FN = "_ZN4ssym18stwo_merkle_kernelILi8ELb1ELi0EEEvNS_10StwoParamsEjNS_6ShaMulE"
out = subprocess.run(["cuobjdump", "-sass", "-fun", FN, os.path.join(ROOT, "stark-symphony_b200", "libssym.so")], capture_output=True, text=True).stdout
lines = [re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip() for l in out.splitlines() if re.match(r"^\s+/\*[0-9a-f]{4}\*/", l)]
addr = lambda l: int(l[2:6], 16)


def op(l):
    t = l.split("*/", 1)[1].split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


back = [(addr(l), int(re.search(r"0x([0-9a-f]+)", l.split("BRA.U UP0,")[1]).group(1), 16)) for l in lines if "BRA.U UP0," in l]
(d_end, d_start), (p_end, p_start) = back[0], back[1]  # data-block loop, padding-block loop
first_k = next(addr(l) for l in lines if d_start <= addr(l) < d_end and "LDCU.64" in l and "c[0x3]" in l)
round_start = max(addr(l) for l in lines if addr(l) < first_k and op(l) in ("LDC", "BRA"))  # the rounds start right after the schedule's exit
sched_start = next(addr(l) for l in lines if addr(l) > d_start and "BRA.U !UP0" in l) + 0x10
secs = [("data block, message schedule: 16 words (skipped in group 0: x3 per compression)", sched_start, round_start),
        ("data block, 16 rounds (x4 per compression)", round_start, d_end + 0x10),
        ("padding block, 16 rounds with K+W from `c_sha_kwpad4` (x4 per compression)", p_start, p_end + 0x10)]
rows = []
for name, a, b in secs:
    s = [l for l in lines if a <= addr(l) < b]
    c = {"SHF": 0, "LOP3": 0, "IMAD": 0}
    ur = 0
    for l in s:
        o = op(l)
        if o in c:
            c[o] += 1
        if o == "IMAD" and "UR" in l.split("IMAD", 1)[1].split(",")[2]:
            ur += 1
    rows.append((name, f"`0x{a:04x}..0x{b:04x}`", len(s), c["SHF"], c["LOP3"], c["IMAD"], ur, len(s) - sum(c.values())))
alu = (rows[0][3] + rows[0][4]) * 3 + (rows[1][3] + rows[1][4]) * 4 + (rows[2][3] + rows[2][4]) * 4
fma = rows[0][5] * 3 + rows[1][5] * 4 + rows[2][5] * 4
print(f"""# K3 `stwo_merkle_kernel<ADDMODE=8, ROLLED=1>` — SASS excerpt (final round-2 build)

`cuobjdump -sass stark-symphony_b200/libssym.so` (sm_100a cubin of `csrc/stwo_kernels.cu`), function `{FN}`: {len(lines)} instructions, one hashing loop (one
iteration = one step of a chain = one 64-byte message) that contains the two rolled loops below (4 iterations each per compression).  No `HMMA` /
`UTC*MMA` / `UTMALDG`: nothing here is a contraction, and the 32-byte siblings are read with `LDG.E.128` straight into registers one level ahead of
their use (DESIGN.md section 4).  This file: `profiles/make_k3_sass_excerpt.py`.

| section | address range | instructions | SHF (ALU pipe) | LOP3 (ALU pipe) | IMAD (FMA pipe) | of which with the multiplier in a uniform register | other (uniform / control / LDCU of K) |
|---|---|---|---|---|---|---|---|""")
for r in rows:
    print("| " + " | ".join(str(x) for x in r) + " |")
print(f"""
Per pair hash (sha256_pair = one data-block compression + the constant padding-block compression): {alu} ALU-pipe instructions (SHF + LOP3) and {fma}
FMA-pipe instructions (the additions, as `IMAD x, one, y` with an opaque multiplier).  ncu counts 864 ALU-pipe lane-instructions per compression
(profiles/step_pipe_counts.json: 104.0 M warp instructions per 1024 x 3760 compressions) against {alu // 2} from this table — the difference is the leaf hashes, the
select / compare code around the loop and the loop control.

Until round 2 the build kept the multiplier `1` of the round additions in a VECTOR register (`IMAD R34, R35, R32, R34`: three vector-register operands;
ncu: 0.62 dispatch stalls per issued instruction).  An `IMAD` takes one operand through the uniform datapath, and the one addition per round whose addend is
K[t] from the constant bank needs that slot for K — so that addition now has an opaque `1` of its own in a vector register (kernel argument `onek`), and the
others read theirs per 16-round group from `c_sha_ones[grp]` (`LDCU`), which ptxas keeps uniform: dispatch stalls 0.16 per issue, ALU pipe 82.7 -> 88.9 %
of active cycles, 0.243 -> 0.229 ms per launch (0.223 with one warp per CTA).

## The first rounds of the data-block round section

```""")
for l in [l for l in lines if round_start <= addr(l) < round_start + 0x10 * 56]:
    print(l)
print("```\n\n## The message-schedule section (first 24 instructions)\n\n```")
for l in [l for l in lines if sched_start <= addr(l) < sched_start + 0x10 * 24]:
    print(l)
print("""```

`SHF.R.W` = the rotations of Sigma0 / Sigma1 / sigma0 / sigma1 (`SHF.R.U32.HI RZ` = the plain shifts of sigma0 / sigma1), `LOP3.LUT ... 0x96` = a
three-way xor, `0xb8` / `0xe8` = Ch / Maj in one instruction each, `IMAD Rd, Ra, URx, Rb` = a 32-bit addition issued on the FMA pipe (`URx` = 1, opaque to
ptxas), `IMAD Rd, Rw, Rone, UR6|UR7` = W[t] + K[t] with K from `LDCU.64 c[0x3][...]`.""")
