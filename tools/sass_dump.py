"""Normalised SASS (instruction text, no addresses / encodings) of every kernel in a .so, one file per kernel:
    python tools/sass_dump.py drawingspinup_b200/lib/libdsu_b200.so /tmp/sass_a ; (edit, rebuild) ; ... /tmp/sass_b ; cmp the files.
Used to prove that adding an experimental template path leaves the default instantiations byte-identical."""
import sys,re,os,subprocess
so,outdir=sys.argv[1],sys.argv[2]
os.makedirs(outdir,exist_ok=True)
txt=subprocess.run(['cuobjdump','-sass',so],capture_output=True,text=True).stdout
cur=None; out={}
for line in txt.splitlines():
    m=re.search(r'Function : (\S+)',line)
    if m: cur=m.group(1); out[cur]=[]; continue
    if cur is None: continue
    m=re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);',line)
    if m: out[cur].append(m.group(1).strip())
for k,v in out.items():
    open(os.path.join(outdir,k[:80]+'.txt'),'w').write('\n'.join(v))
    print(len(v),k[:70])
