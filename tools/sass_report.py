#!/usr/bin/env python3
"""SASS summary of every kernel of libt2b200.so's objects: registers, opcode histogram, the instructions that show which
hardware path a kernel uses (DPX s16x2 min/max, PRMT, bulk / async copies, warp shuffles, barriers).
usage: python tools/sass_report.py > profiles/rNN_sass.txt     (needs the objects under sdr_receiver_dvb_t2_b200/build/)"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
NOTABLE = ('UBLKCP', 'LDGSTS', 'VIADDMNMX', 'VIMNMX', 'PRMT', 'SHFL', 'BAR', 'MEMBAR', 'FENCE', 'ATOM', 'RED', 'LDS', 'STS', 'LDG', 'STG',
           'IMAD', 'LOP3', 'FFMA', 'FADD', 'FMUL', 'MUFU', 'LDC', 'NANOSLEEP', 'VOTE', 'MATCH')


def main():
    for obj in sorted(glob.glob(os.path.join(ROOT, 'sdr_receiver_dvb_t2_b200', 'build', '*.o'))):
        out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
        res = subprocess.run(['cuobjdump', '-res-usage', obj], capture_output=True, text=True).stdout
        regs = dict(re.findall(r'Function (\S+):\s*\n\s*REG:(\d+)', res))
        for m in re.finditer(r'Function : (\S+)\n(.*?)(?=\n\s*Function :|\Z)', out, re.S):
            name, body = m.group(1), m.group(2)
            ops = collections.Counter()
            for line in body.splitlines():
                mm = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
                if mm:
                    ops[mm.group(1)] += 1
            if not ops:
                continue
            dem = subprocess.run(['cu++filt', name], capture_output=True, text=True).stdout.strip() or name
            dem = re.sub(r'\(anonymous namespace\)::', '', dem)
            total = sum(ops.values())
            print('== %s  [%s]' % (dem[:150], os.path.basename(obj)))
            print('   %d instructions, %s registers' % (total, regs.get(name, '?')))
            fam = collections.Counter()
            for k, v in ops.items():
                fam[k.split('.')[0]] += v
            print('   top: ' + ', '.join('%s %d' % kv for kv in fam.most_common(12)))
            notable = {k: v for k, v in ops.items() if any(k.startswith(n) for n in ('UBLKCP', 'LDGSTS', 'VIADDMNMX', 'VIMNMX', 'SHFL', 'BAR', 'FENCE', 'ATOM', 'NANOSLEEP', 'UTMA', 'SYNCS'))}
            if notable:
                print('   notable: ' + ', '.join('%s x%d' % kv for kv in sorted(notable.items())))
            print()


if __name__ == '__main__':
    main()
