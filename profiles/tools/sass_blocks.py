"""Summarise a `cuobjdump -sass -fun <kernel>` dump by basic block: instruction count and the fp64 / shared / global /
control mix, with loop back-edges marked.  Used to check loop bodies offline (no GPU) before spending GPU time.
    python profiles/tools_sass_blocks.py dump.txt [min_instr]"""
import re
import sys

lines = open(sys.argv[1]).read().splitlines()
minn = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ins = re.compile(r'^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);')
prog = []
for ln in lines:
    m = ins.match(ln)
    if m:
        prog.append((int(m.group(1), 16), m.group(2).strip()))
targets = set()
for a, t in prog:
    if re.search(r'\b(BRA|BSSY|CALL|JMP)', t):
        m = re.search(r'0x([0-9a-f]+)\s*$', t)
        if m:
            targets.add(int(m.group(1), 16))
blocks, cur = [], []
for a, t in prog:
    if a in targets and cur:
        blocks.append(cur)
        cur = []
    cur.append((a, t))
    if re.search(r'\b(BRA|EXIT|RET|BRX|JMP)\b', t):
        blocks.append(cur)
        cur = []
if cur:
    blocks.append(cur)
for b in blocks:
    if len(b) < minn:
        continue
    txt = [t for _, t in b]
    def cnt(pat):
        return sum(1 for i in txt if re.search(pat, i))
    last = txt[-1]
    back = ''
    m = re.search(r'BRA.*0x([0-9a-f]+)\s*$', last)
    if m and int(m.group(1), 16) <= b[0][0]:
        back = ' <-- back edge to %04x' % int(m.group(1), 16)
    print(f'{b[0][0]:05x}-{b[-1][0]:05x} n={len(b):4d} f64={cnt(r"^(@!?U?P\d+ )?D(FMA|MUL|ADD|SETP)"):3d} mufu={cnt("MUFU"):2d} cvt={cnt(r"F2I|I2F|FRND|F2F"):2d} lds={cnt(r"LDS"):2d} '
          f'sts={cnt(r"STS"):2d} ldg={cnt(r"LDG|LD\.E"):2d} stg={cnt(r"STG|ST\.E"):2d} ldl={cnt(r"LDL|STL"):2d} vote={cnt("VOTE"):1d} sel={cnt(r"SEL"):3d} '
          f'int={cnt(r"I(MAD|ADD3|MNMX)|LEA|LOP3|SHF|VIADD|ISETP"):3d} | {last[:50]}{back}')
print('total instructions', len(prog))
