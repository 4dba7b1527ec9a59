#!/bin/bash
# SASS of the shipped library's hot kernels -> profiles/r2_k_icp_loop.sass.gz, profiles/r2_k_knn.sass.gz (+ opcode counts)
set -e
SO=mimosa_b200/lib/libmimosa_b200.so
cuobjdump -sass $SO > /tmp/all.sass
python - <<'PY'
import gzip, re, collections
txt = open("/tmp/all.sass").read()
funcs = re.split(r"\n\s*Function : ", txt)
want = {"k_icp_loopILi5ELi19": "profiles/r2_k_icp_loop.sass.gz", "k_knnILi5": "profiles/r2_k_knn.sass.gz"}
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    for key, out in want.items():
        if key in name:
            body = "Function : " + f
            lines = [l for l in body.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l)]
            ops = collections.Counter(re.sub(r"^(@!?U?P\d+\s+)?", "", l.split("*/", 1)[1].strip()).split(" ")[0].split(".")[0].rstrip(";") for l in lines)
            head = "# %s: %d SASS instructions; most frequent opcodes: %s\n" % (name, len(lines), ", ".join("%s %d" % kv for kv in ops.most_common(14)))
            special = {k: v for k, v in ops.items() if k in ("DMMA", "UBLKCP", "UTMALDG", "SYNCS", "UTCMMA", "HMMA", "DMMA", "ACQBULK", "PREEXIT", "BAR", "ATOMG", "REDG", "MATCH", "REDUX")}
            head += "# barrier / atomic / bulk-copy / tensor opcodes present: %s\n" % (special or "none")
            gzip.open(out, "wt").write(head + body)
            print(out, head.strip())
PY
