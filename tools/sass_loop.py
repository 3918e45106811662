"""Instruction mix of the main loop of a tile kernel, from the SASS of an object file or library.

    python tools/sass_loop.py build/obj/fwd_gelu.o 'GeluFn, __nv_bfloat16, 3' [--elems 32] [--dump]

The main loop is taken to be the innermost backward branch that spans the first 128-bit global
load of the kernel; the block-uniform slow path (CALL ... search_exact) is left out of the count
when it sits inside the loop.  `--elems` = elements per lane per loop iteration (U * 8).
"""
import argparse
import collections
import re
import subprocess

ALU = {'FSEL', 'FSETP', 'LOP3', 'SEL', 'IADD3', 'SHF', 'PRMT', 'ISETP', 'LEA', 'FMNMX', 'IADD', 'VIADD', 'PLOP3',
       'I2FP', 'F2FP', 'MOV', 'VIMNMX', 'VIMNMX3', 'FSET', 'P2R', 'R2P', 'BMSK', 'SGXT', 'FLO', 'POPC', 'HSETP2', 'HSET2'}
FMA = {'FFMA', 'FMUL', 'FADD', 'IMAD', 'HFMA2', 'FFMA2', 'FMUL2', 'FADD2', 'HMUL2', 'HADD2'}
MEM = {'LDG', 'STG', 'LDS', 'STS', 'SHFL', 'LDSM', 'ATOMS', 'LDC', 'ULDC'}
CTL = {'BRA', 'BSSY', 'BSYNC', 'WARPSYNC', 'CALL', 'RET', 'EXIT', 'NOP', 'BAR', 'S2R', 'S2UR', 'R2UR', 'CS2R', 'VOTE', 'VOTEU', 'BREAK'}


def opcode(ins):
    return re.sub(r'^@!?U?P\d+\s+', '', ins).split()[0].split('.')[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('lib')
    ap.add_argument('pattern', help='substring of the demangled kernel name')
    ap.add_argument('--kernel', default='forward_tiles_kernel')
    ap.add_argument('--elems', type=int, default=32)
    ap.add_argument('--dump', action='store_true')
    args = ap.parse_args()
    out = subprocess.run(f'cuobjdump -sass {args.lib} | c++filt', shell=True, capture_output=True, text=True).stdout
    body, on = [], False
    for line in out.splitlines():
        if 'Function :' in line:
            on = args.kernel in line and args.pattern in line
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if on and m:
            body.append((int(m.group(1), 16), m.group(2).strip()))
    if not body:
        raise SystemExit('kernel not found')
    addr = {a: i for i, (a, _) in enumerate(body)}
    # every innermost backward branch that spans a 128-bit global load is a candidate; the main
    # loop is the largest one that does not call out of line (the exact-search loop does)
    loads = [i for i, (_, ins) in enumerate(body) if re.search(r'LDG\.E(\.NA)?\.128|LDGSTS', ins)]
    loops = {}
    for first in loads:
        best = None
        for i, (a, ins) in enumerate(body):
            m = re.search(r'\bBRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)', ins)
            if m and i > first:
                t = addr.get(int(m.group(1), 16))
                if t is not None and t <= first and (best is None or i - t < best[1] - best[0]):
                    best = (t, i)
        if best:
            loops[best] = sum(1 for _, ins in body[best[0]:best[1] + 1] if opcode(ins) == 'CALL')
    if not loops:
        raise SystemExit('no loop around a 128-bit load')
    clean = [k for k, calls in loops.items() if calls == 0] or list(loops)
    loop = max(clean, key=lambda k: k[1] - k[0])
    insns = body[loop[0]:loop[1] + 1]
    # drop the slow path: any stretch that CALLs out of line, between the branch that enters it and its join
    keep, skip_until = [], -1
    for i, (a, ins) in enumerate(insns):
        keep.append((a, ins))
    calls = [i for i, (_, ins) in enumerate(keep) if opcode(ins) == 'CALL']
    slow = set()
    if calls:
        # slow-path blocks are delimited by the nearest preceding conditional BRA and the next BRA/BSYNC
        blocks, start = [], None
        for i, (_, ins) in enumerate(keep):
            pass
        i = 0
        while i < len(calls):
            lo = calls[i]
            while lo > 0 and opcode(keep[lo - 1][1]) != 'BRA':
                lo -= 1
            hi = calls[i]
            while hi + 1 < len(keep) and opcode(keep[hi][1]) not in ('BRA', 'BSYNC'):
                hi += 1
            slow.update(range(lo, hi + 1))
            while i < len(calls) and calls[i] <= hi:
                i += 1
    fast = [x for i, x in enumerate(keep) if i not in slow]
    c = collections.Counter(opcode(ins) for _, ins in fast)
    n = args.elems
    cat = lambda names: sum(v for k, v in c.items() if k in names) / n  # noqa: E731
    other = sum(v for k, v in c.items() if k not in ALU | FMA | MEM | CTL | {'MUFU'}) / n
    print(f'{args.pattern}: loop {len(fast)} instr (+{len(slow)} slow-path) = {len(fast) / n:.2f}/elem;  '
          f'ALU {cat(ALU):.2f}  FMA {cat(FMA):.2f}  MUFU {c["MUFU"] / n:.2f}  MEM {cat(MEM):.2f}  CTL {cat(CTL):.2f}  other {other:.2f}')
    print('   ', dict(c.most_common(30)))
    if args.dump:
        for i, (a, ins) in enumerate(keep):
            print(f'{a:05x} {"~" if i in slow else " "} {ins}')


if __name__ == '__main__':
    main()
