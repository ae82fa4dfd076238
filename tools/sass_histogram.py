"""SASS opcode histogram per kernel of gaboflow_b200/libgabo_b200.so (cuobjdump -sass), the evidence for which
hardware paths each kernel uses: DMMA (FP64 tensor), UTCHMMA/UTCQMMA + LDTM/STTM (tcgen05 + TMEM), UTMALDG/UTMASTG (TMA
tensor copies), UBLKCP (TMA bulk copies), SYNCS (mbarrier), LDGSTS (cp.async), LDL/STL (spills).
Usage: python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'gaboflow_b200', 'libgabo_b200.so')
KEYS = ['DMMA', 'DFMA', 'DADD', 'DMUL', 'UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'UTCBAR', 'UTCCP', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG',
        'UBLKCP', 'SYNCS', 'LDGSTS', 'HMMA', 'FFMA', 'FFMA2', 'MUFU', 'LDL', 'STL', 'LDS', 'STS', 'LDG', 'STG', 'SHFL',
        'BAR', 'WARPSYNC', 'ATOM', 'RED', 'MEMBAR', 'CCTL']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = collections.Counter()
            kernels[m.group(1)] = cur
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur is not None:
            cur[m.group(1).split('.')[0]] += 1
            cur['_total'] += 1
    names = subprocess.run(['c++filt'], input='\n'.join(kernels), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    print(f'# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels; counts are static instructions')
    for (mangled, c), name in zip(kernels.items(), names):
        name = name.replace('(anonymous namespace)::', '').replace('void ', '')
        name = re.sub(r'\((?:[^()]|\([^()]*\))*\)\s*$', '', name)
        shown = ' '.join(f'{k}={c[k]}' for k in KEYS if c[k])
        print(f'{name}: total={c["_total"]} {shown}')
        total.update(c)
    print('# whole library: ' + ' '.join(f'{k}={total[k]}' for k in KEYS))


if __name__ == '__main__':
    main()
