import sys, json
for l in (open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin):
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    if 'ms' in d: print('  %8.4f ms  %5.1f%%  %s' % (d['ms'], 100 * d['frac_of_hbm_peak'], d['kernel']))
