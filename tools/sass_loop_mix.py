"""Static instruction mix of the loop bodies (backward branches) of one kernel's SASS.
usage: cuobjdump -sass lib.so | python tools/sass_loop_mix.py <mangled-name-fragment>"""
import collections
import re
import sys

frag = sys.argv[1]
on, ins = False, []
for l in sys.stdin:
    if "Function :" in l:
        on = frag in l
        continue
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
print("instructions:", len(ins))
for a, t in ins:
    if "BRA" in t:
        m2 = re.search(r"0x([0-9a-f]+)", t)
        if m2 and int(m2.group(1), 16) < a:
            tgt = int(m2.group(1), 16)
            body = [x for b, x in ins if tgt <= b <= a]
            c = collections.Counter()
            for x in body:
                x = re.sub(r"^@!?U?P\w+\s+", "", x)
                op = x.split()[0]
                base = op.split(".")[0]
                if base == "IMAD" and ".MOV" in op:
                    base = "IMAD.MOV"
                c[base] += 1
            fp64 = sum(c[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
            print(f"loop {hex(tgt)}..{hex(a)}: {len(body)} instructions, {fp64} on the fp64 pipe")
            print("  ", c.most_common(32))
