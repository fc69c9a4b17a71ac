"""The text summaries under profiles/ are regenerated from the ncu artefacts by
tools/ncu_summary.py; this checks the launch-list part against the committed CSV
(the .ncu-rep based parts need ncu and are made after each GPU capture)."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))


import pytest


@pytest.mark.parametrize("csv_name,bench_name", [('r01_launches_c4_s3.csv', 'r01_bench_c4_s3.json'),
                                                  ('r02_launches_final.csv', 'r02_bench_c4_final.json')])
def test_launch_shares_from_committed_csv(tmp_path, csv_name, bench_name):
    import ncu_summary
    src = os.path.join(ROOT, 'profiles', csv_name)
    dst = tmp_path / 'shares.txt'
    ncu_summary.launches(src, str(dst))
    text = dst.read_text()
    tail = text[text.index('shares among the exposure kernels'):]
    shares = {m.group(1): float(m.group(2)) for m in re.finditer(r'(k_\w+)[^\n]*?\s([0-9.]+)%', tail)}
    assert abs(sum(shares.values()) - 100.0) < 0.5
    # the electron thrower dominates; the per-pixel pass and the count sampler share the rest
    assert shares['k_throw_philox'] > 70 and shares['k_reads_native'] < 15 and shares['k_counts_window'] < 15
    # and it agrees with the live CUDA-event stage times of the committed bench line
    import json
    line = json.loads(open(os.path.join(ROOT, 'profiles', bench_name)).read().strip().splitlines()[-1])
    st = {k: v for k, v in line['stage_ms'].items() if k.startswith('k_')}
    live = 100.0 * st['k_throw'] / (st['k_throw'] + st['k_reads'] + st['k_counts'])
    assert abs(live - shares['k_throw_philox']) < 3.0
