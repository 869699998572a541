"""Scoreboard assignment of a kernel, decoded from the control words of its SASS (cuobjdump; no GPU needed): for every
instruction that sets a write barrier, waits on one or branches, the barrier it sets (W), the read barrier (R) and
the wait mask.  This is how the register ring of the camera half was understood: the loads of all
stages carried W5 and the first use of every stage waited on bit 5, i.e. for the youngest load.

    python tools/sass_scoreboards.py povar_b200/lib/libpovar_b200.so k_passB_e0_v2ILb0ELb0
"""
import re,sys,subprocess
lib,pat=sys.argv[1],sys.argv[2]
txt=subprocess.run(['cuobjdump','-sass',lib],capture_output=True,text=True).stdout
lines=txt.split('\n')
on=False; out=[]
i=0
while i<len(lines):
    l=lines[i]
    if 'Function :' in l: on = pat in l
    if on:
        m=re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s*/\* 0x([0-9a-f]+) \*/',l)
        if m and i+1<len(lines):
            m2=re.match(r'\s+/\* 0x([0-9a-f]+) \*/',lines[i+1])
            if m2:
                hi=int(m2.group(1),16); ctl=(hi>>41)&0x1fffff
                out.append((m.group(1),m.group(2).strip(),ctl&0xf,(ctl>>5)&7,(ctl>>8)&7,(ctl>>11)&0x3f))
                i+=2; continue
    i+=1
for a,ins,stall,wb,rb,wait in out:
    if wb!=7 or wait or 'BRA' in ins:
        print(a,ins[:62].ljust(62),'W%s'%('-' if wb==7 else wb),'R%s'%('-' if rb==7 else rb),'wait=%s'%format(wait,'06b'))
