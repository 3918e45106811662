"""SASS excerpt of the projection kernel: python tools/sketch_sass.py > profiles/r02_sketch_sass.txt
Counts every Blackwell tensor-core / TMEM / TMA / mbarrier mnemonic in the two sketch_kernel instantiations
(cuobjdump -sass of the built object) and prints the MMA issue sequence of the pair-mode kernel."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OBJ = ROOT / 'build' / 'obj' / 'sketch.o'
WANTED = re.compile(r'^(UTC|LDTM|STTM|UTMA|UBLKCP|SYNCS|ELECT|UCGABAR|CGABAR|ACQBULK|FENCE\.VIEW\.ASYNC|MEMBAR)')


def main():
    text = subprocess.run(['cuobjdump', '-sass', str(OBJ)], capture_output=True, text=True, check=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in text.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            kernels[name] = []
            continue
        m = re.search(r'/\*[0-9a-f]{4,5}\*/\s+(.*?);', line)
        if m and name:
            kernels[name].append(m.group(1).strip())
    print('# SASS excerpt of the projection kernel (fewbit_b200/csrc/sketch.cu), cuobjdump -sass build/obj/sketch.o, nvcc 12.9 sm_100a')
    print('# (tools/sketch_sass.py).  Per sketch_kernel instantiation: counts of every tensor-core / TMEM / TMA / mbarrier mnemonic --')
    print('#   UTCHMMA(.2CTA) = tcgen05.mma (cta_group::1 / ::2), UTCBAR(.2CTA.MULTICAST) = tcgen05.commit, LDTM = tcgen05.ld,')
    print('#   UTMALDG.2D(.2CTA) = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk shared::cluster, SYNCS = mbarrier ops,')
    print('#   UTCATOMSWS / UTCALLOC-family = tcgen05.alloc / dealloc -- then the MMA issue sequence of one 64-token stage.')
    for name, ins in kernels.items():
        if 'sketch_kernel' not in name:
            continue
        print(f'\n## {name}   ({len(ins)} instructions)')
        counts = collections.Counter()
        for text in ins:
            op = re.sub(r'^@!?U?P\d+\s+', '', text).split()[0]
            if WANTED.match(op):
                counts[op] += 1
        print('   ' + ', '.join(f'{k} x{v}' for k, v in sorted(counts.items())))
        mma = [i for i, text in enumerate(ins) if 'UTCHMMA' in text]
        if mma:
            print('   --- MMA issue sequence (4 K-steps x 3 feature blocks = 12 MMAs, then the commits that free the X stage / S slot) ---')
            last = mma[-1]
            while last + 1 < len(ins) and last - mma[-1] < 24 and 'BRA' not in ins[last]:
                last += 1
            for text in ins[max(mma[0] - 6, 0):last + 1]:
                if any(k in text for k in ('UTCHMMA', 'UTCBAR', 'UIADD3', 'UMOV', 'ELECT', 'SYNCS', 'R2UR', 'BRA')):
                    print('   ' + text)


if __name__ == '__main__':
    sys.exit(main())
