"""Aggregate an ncu source page (csv from --page source --print-source cuda,sass) by named line ranges.
usage: ncu_regions.py src.csv"""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; cur=None; lines=[]
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if r[0]=="Line No":
        hdr=r; ie=hdr.index("Instructions Executed"); ss=hdr.index("# Samples"); continue
    if hdr and r[0].isdigit():
        try: lines.append((cur,int(r[0]),int(r[ie-len(hdr)] or 0),int(r[ss-len(hdr)] or 0)))
        except ValueError: pass
def grep_line(path, text):
    for n,l in enumerate(open(path),1):
        if text in l: return n
    raise SystemExit("marker not found: "+text)
G="mimosa_b200/csrc/mb_search_group.cuh"; F="mimosa_b200/csrc/mb_factor.cu"
g=lambda t: grep_line(G,t); f=lambda t: grep_line(F,t)
regions=[
 ("group: setup+groups", "mb_search_group.cuh", g("template <int K, int ROWS>\nMB_DEV GroupLane"[:24]) if False else 1, g("// ---- (2) resolution")),
 ("group: resolution", "mb_search_group.cuh", g("// ---- (2) resolution"), g("// ---- (3) staging")),
 ("group: staging", "mb_search_group.cuh", g("// ---- (3) staging"), g("// ---- per-query set-up that overlaps")),
 ("group: bounds/lambdas(offer)", "mb_search_group.cuh", g("// ---- per-query set-up that overlaps"), g("// ---- (4) the query's own voxel")),
 ("group: own voxel", "mb_search_group.cuh", g("// ---- (4) the query's own voxel"), g("// ---- (5) which neighbours")),
 ("group: todo mask", "mb_search_group.cuh", g("// ---- (5) which neighbours"), g("// ---- (6) the surviving")),
 ("group: neighbour loop", "mb_search_group.cuh", g("// ---- (6) the surviving"), g("// The winners' global indices")),
 ("group: resolve", "mb_search_group.cuh", g("// The winners' global indices"), 100000),
 ("math", "mb_math.cuh", 1, 100000),
 ("search.cuh (hash etc)", "mb_search.cuh", 1, 100000),
 ("factor: fit_plane", "mb_factor.cu", f("template <int K>\n"[:16]) if False else 100, f("// Sum `count` rows")),
 ("factor: prologue", "mb_factor.cu", f("k_linearize(MapView mv"), f("const size_t n_tiles = (fv.n + 31) / 32;")),
 ("factor: A", "mb_factor.cu", f("const size_t n_tiles = (fv.n + 31) / 32;"), f("// ---- B: search + plane fit")),
 ("factor: B glue", "mb_factor.cu", f("// ---- B: search + plane fit"), f("// ---- C: residual, Jacobian")),
 ("factor: C", "mb_factor.cu", f("// ---- C: residual, Jacobian"), f("// ---- block partial -> group partial")),
 ("factor: reduction+finalize", "mb_factor.cu", f("// ---- block partial -> group partial"), f("// Component localizabilities")),
]
ti=sum(x[2] for x in lines); ts=sum(x[3] for x in lines)
acc={r[0]:[0,0] for r in regions}; other=[0,0]; of={}
for fn,n,i,s in lines:
    for name,file,a,b in regions:
        if fn==file and a<=n<b:
            acc[name][0]+=i; acc[name][1]+=s; break
    else:
        other[0]+=i; other[1]+=s; of[fn]=of.get(fn,0)+i
print("total warp-instr %d  samples %d"%(ti,ts))
for name,_,_,_ in regions:
    print("%-32s %5.1f%% instr (%9d)  %5.1f%% samples"%(name,100*acc[name][0]/ti,acc[name][0],100*acc[name][1]/max(ts,1)))
print("%-32s %5.1f%% instr  %5.1f%% samples  %s"%("other",100*other[0]/ti,100*other[1]/max(ts,1),sorted(of.items(),key=lambda kv:-kv[1])[:5]))
