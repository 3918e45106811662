"""Count the SASS instruction mix of the main loop of a kernel (fast path only: stops at the
first divergent slow-path branch target is not attempted; reports whole loop body)."""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
elems = int(sys.argv[3]) if len(sys.argv) > 3 else 16
out = subprocess.run(f"cuobjdump -sass {lib} | c++filt", shell=True, capture_output=True, text=True).stdout
body, on = [], False
for line in out.splitlines():
    if 'Function :' in line:
        on = 'forward_tiles_kernel' in line and pat in line or ('backward_tiles_kernel' in line and pat in line)
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if on and m:
        body.append((int(m.group(1), 16), m.group(2).strip()))
first = next(i for i, (_, ins) in enumerate(body) if 'LDG.E.NA.128' in ins or 'LDG.E.128' in ins)
last = next(i for i, (_, ins) in enumerate(body) if i > first and re.search(r'\bEXIT\b', ins))
loop = body[first:last]
def op(ins):
    return re.sub(r'^@!?U?P\d+\s+', '', ins).split()[0].split('.')[0]
c = collections.Counter(op(i) for _, i in loop)
alu = {'FSEL','FSETP','LOP3','SEL','IADD3','SHF','PRMT','ISETP','LEA','FMNMX','IADD','VIADD','PLOP3','I2FP','F2FP','MOV','BRA','BSSY','BSYNC','SHFL','VOTE','WARPSYNC'}
fma = {'FFMA','FMUL','FADD','IMAD','HFMA2'}
print(f'{pat}: loop {len(loop)} instr = {len(loop)/elems:.1f}/elem;  ALU-ish {sum(v for k,v in c.items() if k in alu)/elems:.1f}  FMA-ish {sum(v for k,v in c.items() if k in fma)/elems:.1f}  MUFU {c["MUFU"]/elems:.1f}  LDS {c["LDS"]/elems:.1f}')
print('   ', dict(c.most_common(18)))
if len(sys.argv) > 4:
    for a, i in loop[:int(sys.argv[4])]:
        print(f'{a:05x} {i}')
