#!/usr/bin/env python
"""Per-opcode SASS histogram of the K1 step loop (rk4_rollout_kernel<double, front-steer, fast, sliced, tabulated>).

    python tools/k1_sass_hist.py python_motionplanning_b200/_build/rollout_kernels_f64.o [--full]

The step loop is the innermost backward branch that contains the 16 MUFU / >= 500 FP64 instructions of one RK4 step;
the out-of-line checked-step call sequence inside it (between the guarding BRA and the CALL's return) is cold and is
excluded.  Prints FP64 / other counts and the cost-model cycles 2.17 x FP64 + other."""
import collections
import re
import subprocess
import sys

FUN = "_ZN6b200mp18rk4_rollout_kernelIdLb1ELb0ELb0ELb1ELb1ELb0ELb0EEEvNS_10RolloutDevIT_EENS_9DevParamsIS2_EENS_10SliceSchedE"
FP64 = {"DFMA", "DMUL", "DADD", "DSETP"}


def main():
    obj = sys.argv[1]
    global FUN
    f32 = "--f32" in sys.argv          # the FP32 twin (K1f) in rollout_kernels_f32.o
    if f32:
        FUN = FUN.replace("kernelId", "kernelIf")
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", FUN, obj], capture_output=True, text=True).stdout
    ins = []
    for ln in txt.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            t = m.group(2).strip()
            op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
            ins.append((int(m.group(1), 16), op, t))
    loops = []
    for a, op, t in ins:
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                loops.append((int(m.group(1), 16), a))
    best = None
    for lo, hi in loops:
        body = [i for i in ins if lo <= i[0] <= hi]
        fp = sum(1 for i in body if i[1].split(".")[0] in FP64)
        if f32:
            fp = sum(1 for i in body if i[1].split(".")[0] in ("FFMA", "FMUL", "FADD"))
        if fp >= 400 and (best is None or hi - lo < best[1] - best[0]):
            best = (lo, hi)
    lo, hi = best
    body = [i for i in ins if lo <= i[0] <= hi]
    # cold region: from the instruction after the forward BRA that skips the fallback to the target of that BRA
    cold = set()
    calls = [i[0] for i in body if i[1].startswith("CALL")]
    for a, op, t in body:
        if op.startswith("BRA") and calls:
            m = re.search(r"0x([0-9a-f]+)", t)
            tgt = int(m.group(1), 16) if m else 0
            if tgt > a and any(a < c < tgt for c in calls):
                cold.update(x[0] for x in body if a < x[0] < tgt)
    hot = [i for i in body if i[0] not in cold]
    cnt = collections.Counter(i[1].split(".")[0] for i in hot)
    fp = sum(cnt[k] for k in FP64)
    other = len(hot) - fp
    print(f"step loop 0x{lo:x}-0x{hi:x}: {len(hot)} hot instructions ({len(cold)} cold excluded): FP64 {fp}, other {other}, "
          f"cost model 2.17*FP64 + other = {2.17 * fp + other:.0f} cycles per warp-step")
    print("  " + ", ".join(f"{k} {v}" for k, v in cnt.most_common()))
    if "--full" in sys.argv:
        full = collections.Counter(i[1] for i in hot)
        print("  " + ", ".join(f"{k} {v}" for k, v in full.most_common()))
    regs = subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True).stdout
    m = re.search(re.escape(FUN) + r".*?\n\s*(REG:\d+.*)", regs)
    if m:
        print("  " + m.group(1).strip())


if __name__ == "__main__":
    main()
