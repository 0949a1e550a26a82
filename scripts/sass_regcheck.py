"""Check that no code reachable after a USETMAXREG uses a register above the budget it set.

    cuobjdump -sass -fun <kernel> file.o > sass.txt ; python scripts/sass_regcheck.py sass.txt

Walks the control-flow graph (BRA targets + fall-through) from every USETMAXREG until EXIT or the next
USETMAXREG and reports the highest general register index touched (vector / 64-bit operands widen it).
"""
import re
import sys

ins = []  # (addr, text)
for l in open(sys.argv[1]):
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr2i = {a: i for i, (a, _) in enumerate(ins)}


def width(txt):
    if '.128' in txt or 'F32x4' in txt:
        return 4
    if '.64' in txt or 'WIDE' in txt or 'DADD' in txt or 'DMUL' in txt or 'DFMA' in txt or 'CS2R' in txt:
        return 2
    return 1


def maxreg(txt):
    mx = -1
    w = width(txt)
    if txt.startswith('UTC') or 'TCGEN' in txt or 'LDTM' in txt:
        m = re.search(r'x(\d+)', txt)
        if m:
            w = max(w, int(m.group(1)))
    for m in re.finditer(r'\bR(\d+)\b', txt):
        mx = max(mx, int(m.group(1)) + w - 1)
    return mx


def succ(i):
    a, t = ins[i]
    out = []
    pred = t.startswith('@')
    body = t.split(' ', 1)[1] if pred else t
    op = body.split(' ')[0]
    if op.startswith('EXIT') and not pred:
        return []
    if op.startswith('BRA') or op.startswith('JMP'):
        m = re.search(r'0x([0-9a-f]+)\s*$', body)
        if m and int(m.group(1), 16) in addr2i:
            out.append(addr2i[int(m.group(1), 16)])
        uncond = not pred and not ('.U' in op or '.DIV' in op or ',' in body and re.search(r'!?U?P\d', body.split(',')[0]))
        if not uncond and i + 1 < len(ins):
            out.append(i + 1)
        if uncond:
            return out
        return out
    if op.startswith('BSSY') or op.startswith('BSYNC') or op.startswith('WARPSYNC') and 'COLLECTIVE' in op:
        m = re.search(r'0x([0-9a-f]+)\s*$', body)
        if m and int(m.group(1), 16) in addr2i:
            out.append(addr2i[int(m.group(1), 16)])
    if i + 1 < len(ins):
        out.append(i + 1)
    return out


marks = [i for i, (_, t) in enumerate(ins) if 'USETMAXREG' in t]
bad = 0
for s in marks:
    budget = int(re.search(r'0x([0-9a-f]+)', ins[s][1].split('CTAPOOL')[1]).group(1), 16)
    seen = set()
    stack = [s + 1]
    mx, where = -1, None
    while stack:
        i = stack.pop()
        if i in seen or i >= len(ins):
            continue
        seen.add(i)
        if 'USETMAXREG' in ins[i][1]:
            continue
        r = maxreg(ins[i][1])
        if r > mx:
            mx, where = r, ins[i]
        stack.extend(succ(i))
    ok = mx < budget
    bad += not ok
    print(f"{ins[s][0]:#x} {ins[s][1]:45s} reachable={len(seen):6d} max R{mx} budget {budget} {'OK' if ok else 'VIOLATION'} at {where[0]:#x} {where[1][:70]}")
sys.exit(1 if bad else 0)
