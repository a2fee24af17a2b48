"""Decode the scheduling control bits of sm_100a SASS as printed by `cuobjdump -sass`.

usage:  cuobjdump -sass -fun '<mangled kernel>' slrbs_b200/csrc/build/pgs_pipe.o > k.sass; python tools/sass_ctrl.py k.sass

Every 128-bit instruction is printed on two lines; the second 64-bit word carries, above the operand bits, the stall count
(bits 41-44), yield (45), the scoreboard the instruction SETS when its result arrives (write barrier, 46-48; 7 = none), the
scoreboard set when its operands have been read (49-51), and the mask of scoreboards it WAITS for (52-57). A scoreboard is a
counter: waiting on one waits for every instruction that was put on it. That is how profiles/README.md established that the
solve of chunk k waited for the loads of chunk k+1 (both on scoreboard 5) and that the fix removed that wait.
Output per instruction:  address  w=<wait mask>  wr=<write barrier>  rd=<read barrier>  st=<stall>  text
"""
import re
import sys
lines=open(sys.argv[1]).read().split('\n')
out=[]
i=0
pat=re.compile(r'^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/')
pat2=re.compile(r'^\s+/\* 0x([0-9a-f]{16}) \*/')
while i<len(lines):
    m=pat.match(lines[i])
    if m and i+1<len(lines):
        m2=pat2.match(lines[i+1])
        if m2:
            hi=int(m2.group(1),16)
            stall=(hi>>41)&0xf; y=(hi>>45)&1; wr=(hi>>46)&7; rd=(hi>>49)&7; wait=(hi>>52)&0x3f
            out.append((m.group(1), m.group(2).strip(), stall,y,wr,rd,wait))
            i+=2; continue
    i+=1
for a,t,stall,y,wr,rd,wait in out:
    print(f"{a} w={wait:06b} wr={'-' if wr==7 else wr} rd={'-' if rd==7 else rd} st={stall:2d} {t[:90]}")
