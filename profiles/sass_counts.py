#!/usr/bin/env python
"""SASS mnemonic counts per kernel:  cuobjdump -sass relation_autoencoder_b200/librae.so | python profiles/sass_counts.py"""
import sys, re, subprocess, collections
cur = None
cnt = collections.defaultdict(collections.Counter)
pat = {'UTCHMMA': r'\bUTCHMMA\b', 'LDTM': r'\bLDTM', 'STTM': r'\bSTTM', 'UBLKCP': r'\bUBLKCP', 'UTCBAR': r'\bUTCBAR', 'SYNCS': r'\bSYNCS',
       'MATCH': r'\bMATCH', 'ATOMS': r'\bATOMS', 'ATOMG/RED': r'\bATOMG|\bRED\b', 'HMMA': r'\bHMMA\b', 'ACQBULK': r'\bACQBULK'}
for line in sys.stdin:
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        cnt[cur]
        continue
    if cur is None:
        continue
    for k, p in pat.items():
        if re.search(p, line):
            cnt[cur][k] += 1
names = list(cnt)
dem = subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.splitlines()
print("SASS mnemonic counts per kernel of relation_autoencoder_b200/librae.so (cuobjdump -sass, sm_100a)")
print("UTCHMMA = tcgen05.mma (kind::f16 / kind::tf32), LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit,")
print("SYNCS = mbarrier ops, MATCH = match.any (own radix sort), ACQBULK = griddepcontrol.wait (programmatic dependent launch),")
print("ATOMS = shared-memory integer atomics (sort histogram), ATOMG/RED = global atomics (integer max / counts only)")
for n, d in sorted(zip(names, dem), key=lambda x: x[1]):
    c = cnt[n]
    if not c:
        continue
    d = d.replace('(anonymous namespace)::', '').replace('void ', '')
    d = re.sub(r'\(.*', '', d).replace('rae::', '')
    print("%-50s %s" % (d[:50], '  '.join('%s=%d' % (k, v) for k, v in sorted(c.items()))))
