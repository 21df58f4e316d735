"""Compare two `cuobjdump -sass` dumps kernel by kernel (instruction text + encodings, whitespace-insensitive).

    cuobjdump -sass old/libb200pt.so > old.sass; cuobjdump -sass mitsuba3_b200/lib/libb200pt.so > new.sass
    python tools/sass_compare.py old.sass new.sass

Used to show that staging an experimental, default-off variant (a new template instantiation) leaves the
device code of every shipped kernel untouched when no GPU is at hand to re-measure.
"""
import re,hashlib,sys
def funcs(path):
    out={}; cur=None; buf=[]
    for line in open(path):
        m=re.match(r'\s*Function : (\S+)',line)
        if m:
            if cur: out[cur]=hashlib.md5(''.join(buf).encode()).hexdigest()
            cur=m.group(1); buf=[]
        elif cur:
            t=re.sub(r'\s+',' ',line).strip()
            if t: buf.append(t+'\n')
    if cur: out[cur]=hashlib.md5(''.join(buf).encode()).hexdigest()
    return out
base=funcs(sys.argv[1]); new=funcs(sys.argv[2])
def norm(k):
    """Template parameters added later with a default value (k_trace_dyn<FIRST, SMEM_ALL[, PHASE = 0[, WIDE = false[, ORDERED = false]]]>)
    change the mangled name of the unchanged instantiation; fold them away."""
    if 'k_trace_dyn' in k:
        for tail in ('ELi0ELb0ELb0EEEv', 'ELi0ELb0EEEv', 'ELi0EEEv'):
            k = k.replace(tail, 'EEEv', 1) if tail in k else k
    if 'k_splat_gauss' in k:          # k_splat_gauss<WEIGHTS_ONLY[, FOLD = false]>
        k = k.replace('ILb0ELb0EEEv', 'ILb0EEEv').replace('ILb1ELb0EEEv', 'ILb1EEEv')
    return k
base={norm(k):v for k,v in base.items()}
newn={norm(k):v for k,v in new.items()}
bad=0
for k in base:
    if k not in newn: print('MISSING',k[:80]); bad+=1
    elif base[k]!=newn[k]: print('DIFF',k[:80]); bad+=1
print(len(base),'kernels compared,',bad,'differ;', len(new)-len(base),'new kernels')
