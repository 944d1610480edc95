"""Feeds damaged archives to the host decoder (bce -ds path, no GPU).  Run as a script by
tests/test_host_coders.py in a child process: a crash shows up as its exit status."""
import json, sys, random
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bce_b200 import host
from oracle import oracle
from tests.inputs import small_cases
arcs=[]
for name,data,prim in small_cases():
    if 2 <= len(data) <= 5000:
        arcs.append(bytes(oracle.compress(data)))
print(len(arcs), 'archives', sum(map(len,arcs)), flush=True)
rng=random.Random(int(sys.argv[1]))
ok=err=0
for it in range(int(sys.argv[2])):
    a=bytearray(rng.choice(arcs))
    mode=rng.randrange(4)
    if mode==0 and len(a)>2:
        for _ in range(rng.randrange(1,4)): a[rng.randrange(len(a))]^=1<<rng.randrange(8)
    elif mode==1 and len(a)>4: a=a[:rng.randrange(2,len(a))//2*2]
    elif mode==2: a[rng.randrange(min(len(a),12))]=rng.randrange(256)
    else: a+=bytes(rng.randrange(256) for _ in range(2*rng.randrange(1,8)))
    try:
        out=host.decompress(bytes(a), low_memory=True); ok+=1
    except Exception as e:
        err+=1
print('ok',ok,'err',err)
