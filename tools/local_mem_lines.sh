#!/bin/bash
# list source lines that generate local-memory (LDL/STL) instructions in the SfT kernel
cd /tmp && rm -f sft_cuda*.cubin && cuobjdump -xelf all /root/repo/defslam_b200/csrc/build/sft_cuda.o >/dev/null 2>&1
nvdisasm -g /tmp/sft_cuda*.cubin > /tmp/sass_g.txt 2>/dev/null
python3 - <<'PY'
import re
cur=None; cnt={}
for line in open('/tmp/sass_g.txt'):
    m=re.search(r'//## File "([^"]+)", line (\d+)',line)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    if re.search(r'\b(LDL|STL)(\.|\b)',line):
        cnt[cur]=cnt.get(cur,0)+1
for k,v in sorted(cnt.items(), key=lambda kv:-kv[1])[:25]: print(v,k)
PY
