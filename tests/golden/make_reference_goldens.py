"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN FUNCTION BODIES.

The reference package cannot be imported here (Python-2 syntax elsewhere in the
modules, astropy / matplotlib / pandas imports at module level), but the
functions on the exposure path are plain numpy.  This script reads the reference
sources where they lie (/root/reference, never copied into the repository),
extracts the named functions / classes with ``ast``, and runs them under
Python 3 with two shims that restore their Python-2 meaning:

  * every ``a / b`` becomes ``_py2div(a, b)``: floor division when both operands
    are integers (no module uses ``from __future__ import division``), true
    division otherwise;
  * ``xrange`` = ``range``; ``u`` (astropy.units) = wayne_b200.units, whose
    Quantity does the same float arithmetic (only where a function needs it).

Outputs: tests/golden/reference_goldens.npz, compared against oracle/ and the
host mirror by tests/test_reference_goldens.py.  Run here, where the reference
exists:   python tests/golden/make_reference_goldens.py
"""
import ast
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get('WAYNE_REFERENCE', '/root/reference') + '/wayne'

from wayne_b200 import units as U  # noqa: E402


def _py2div(a, b):
    ints = (int, np.integer)
    if isinstance(a, ints) and isinstance(b, ints) and not isinstance(a, bool):
        return a // b
    return a / b


class _Div(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(
                ast.Call(func=ast.Name(id='_py2div', ctx=ast.Load()), args=[node.left, node.right], keywords=[]),
                node)
        return node


def extract(path, names, ns=None):
    """Execute the top-level definitions ``names`` of a reference module in a
    fresh namespace (numpy as np, the shims) and return it."""
    ns = dict(ns or {})
    ns.setdefault('np', np)
    ns.setdefault('xrange', range)
    ns.setdefault('u', U)
    ns['_py2div'] = _py2div
    src = open(os.path.join(REF, path)).read()
    try:
        tree = ast.parse(src)
    except SyntaxError:
        # Python-2-only syntax elsewhere in the file: cut out the wanted top-level
        # blocks textually (a block runs until the next statement in column 0)
        lines, blocks, i = src.splitlines(), [], 0
        while i < len(lines):
            ln = lines[i]
            head = ln.split('(')[0].split('=')[0].split(':')[0].replace('def ', '').replace('class ', '').strip()
            if ln[:1] not in (' ', '\t', '#', '') and head in names:
                j = i + 1
                while j < len(lines) and (lines[j][:1] in (' ', '\t', '') or lines[j].startswith(')')):
                    j += 1
                blocks.append('\n'.join(lines[i:j]))
                i = j
            else:
                i += 1
        tree = ast.parse('\n\n'.join(blocks))
    keep = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            keep.append(node)
        elif isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id in names for t in node.targets):
            keep.append(node)
    found = {getattr(n, 'name', None) or n.targets[0].id for n in keep}
    missing = set(names) - found
    assert not missing, (path, missing)
    mod = _Div().visit(ast.Module(body=keep, type_ignores=[]))
    ast.fix_missing_locations(mod)
    exec(compile(mod, path, 'exec'), ns)
    return ns


def plain(q):
    """Value of a dimensionless Quantity (or a plain number / array)."""
    if U.is_quantity(q):
        return np.asarray(q.value, dtype=float) * q.unit.scale
    return np.asarray(q, dtype=float)


out = {}

# ---- grism.py: calibration coefficients, trace, flat field ---------------------------------
g = extract('grism.py', ['wavelength_calibration_coeffs', '_SpectrumTrace', 'g141_trace_coeff',
                         'g141_wl_solution', 'g102_trace_coeff', 'g102_wl_solution', 'G141'],
            ns={'WFC3_IR': object, 'os': os, 'fits': None, 'params': None, 'tools': None})
wl = np.linspace(0.8, 1.9, 37)
pts = [(404.497, 457.427), (50.0, 50.0), (612.3, 171.9), (330.02, 930.4)]
rows = []
for name, a, b in (('g141', g['g141_trace_coeff'], g['g141_wl_solution']),
                   ('g102', g['g102_trace_coeff'], g['g102_wl_solution'])):
    for x, y in pts:
        tr = g['_SpectrumTrace'](x, y, a, b)
        m_wl = U.value_in(tr.m_wl, U.micron)
        c_wl = U.value_in(tr.c_wl, U.micron)
        xs = plain(tr.wl_to_x(wl * U.micron))
        ys = plain(tr.wl_to_y(wl * U.micron))
        rows.append(np.concatenate([[x, y, float(tr.m_t), float(tr.c_t), float(tr.m_w), float(tr.c_w),
                                     float(m_wl), float(c_wl)], xs, ys]))
out['trace_wl'] = wl
out['trace_rows'] = np.array(rows)          # 4 g141 rows then 4 g102 rows

tools_ns = extract('tools.py', ['crop_central_box', 'crop_spectrum_ind', 'bin_centers_to_widths',
                                'bin_centers_to_edges', 'detect_orbits'])
rng = np.random.default_rng(20170410)
flat = [(1 + 0.01 * rng.standard_normal((1014, 1014))).astype('>f4')]
flat += [(0.005 * rng.standard_normal((1014, 1014))).astype('>f4') for _ in range(3)]
fake = types.SimpleNamespace(
    flat_f0=flat[0], flat_f1=flat[1], flat_f2=flat[2], flat_f3=flat[3], flat_wmin=9880.0, flat_wmax=17770.0,
    _get_wavelength_calibration_coeffs=lambda x, y: g['wavelength_calibration_coeffs'](
        x, y, g['g141_trace_coeff'], g['g141_wl_solution']))
fake.flat_xs, fake.flat_ys = np.meshgrid(np.arange(1014), np.arange(1014))
get_flat = g['G141'].get_flat_field
g['tools'] = types.SimpleNamespace(crop_central_box=tools_ns['crop_central_box'])
get_flat.__globals__['tools'] = g['tools']
frame = np.zeros((256, 256))
hit_r = rng.integers(60, 110, 400)
hit_c = rng.integers(30, 230, 400)
frame[hit_r, hit_c] = 1
idx = np.where(frame > 0)
ff = get_flat(fake, 404.497, 457.427, 256, idx)
assert ff.shape == (256, 256) and ff.dtype.itemsize == 4
out['flat_seed'] = np.int64(20170410)
out['flat_idx_r'], out['flat_idx_c'] = idx[0].astype(np.int32), idx[1].astype(np.int32)
out['flat_values'] = np.asarray(ff[idx], dtype=np.float64)
out['flat_off_pixel'] = np.float64(ff[0, 0])
full = get_flat(fake, 404.497, 457.427, 256, None)
out['flat_full_sample'] = np.asarray(full[::17, ::19], dtype=np.float64)

# ---- tools.py -------------------------------------------------------------------------------
wlg = np.sort(rng.uniform(0.85, 1.9, 300))
out['tools_wl'] = wlg
out['tools_crop_ind'] = np.array(tools_ns['crop_spectrum_ind'](0.988, 1.777, wlg))
out['tools_widths'] = tools_ns['bin_centers_to_widths'](wlg)
out['tools_edges'] = tools_ns['bin_centers_to_edges'](wlg)
a66 = np.arange(66 * 66, dtype=float).reshape(66, 66)
out['tools_crop_box'] = tools_ns['crop_central_box'](a66, 32)
out['tools_orbits'] = np.array(tools_ns['detect_orbits'](np.array([1.0, 1.01, 1.05, 1.06, 1.12, 1.121])))

# ---- detector.py: non-linearity Newton solve, bias border, buffer arithmetic ---------------
d = extract('detector.py', ['WFC3_IR'], ns={'os': os, 'fits': None, 'params': None, 'tools': None, 'pd': None})
c2 = (6.4e-7 * (1 + 0.05 * rng.standard_normal((1024, 1024)))).astype('>f4')
c3 = (1e-12 * rng.standard_normal((1024, 1024))).astype('>f4')
zero = np.zeros((1024, 1024), '>f4')
det = types.SimpleNamespace(non_linear_c1=zero, non_linear_c2=c2, non_linear_c3=c3, non_linear_c4=zero)
px = rng.uniform(-50, 70000, (96, 96))
out['nl_seed_note'] = np.array('planes drawn after the flat cube from default_rng(20170410)')
out['nl_c2_sample'] = np.asarray(c2[464:560:19, 464:560:19], dtype=np.float64)
out['nl_input'] = px
out['nl_output'] = d['WFC3_IR'].apply_non_linearity(det, px)
out['bias_pixels'] = d['WFC3_IR'].add_bias_pixels(det, np.arange(64 * 64, dtype=float).reshape(64, 64))
out['pixel_array_shapes'] = np.array([d['WFC3_IR'].gen_pixel_array(det, s, ls).shape
                                      for s in (1024, 512, 256) for ls in (True, False)])
out['exp_per_buffer'] = np.array([d['WFC3_IR'].num_exp_per_buffer(det, n, s)
                                  for n, s in ((5, 256), (15, 1024), (16, 64), (3, 512))])

# ---- trend generators -------------------------------------------------------------------------
cr = extract('trend_generators/cosmic_rays.py', ['BaseCosmicGenerator', 'MinMaxPossionCosmicGenerator'])
np.random.seed(1963)
gen = cr['MinMaxPossionCosmicGenerator'](11., 10000, 35000)
out['cosmic_frame_seed1963_t20_s256'] = gen.cosmic_frame(20.0, 256)
out['cosmic_rate_64'] = np.float64(gen._rate_full_frame_to_size(11., 64))

ssv = extract('trend_generators/scan_speed_varations.py', ['SSVSine'])
ymid = 457.4 + np.linspace(0, 165, 400)
dur = np.full(400, 10.0)
dur[57] = 8.0
out['ssv_y'] = ymid
out['ssv_dur'] = dur
out['ssv_out'] = ssv['SSVSine'](1.5, 1.1, 0).get_subsample_exposure_times(ymid, dur, None, None)
out['ssv_out_phase'] = ssv['SSVSine'](2.5, 0.7, 1.3).get_subsample_exposure_times(ymid, dur, None, None)

# SSVModulatedSine (scan_speed_varations.py:63-171): stochastic; numpy's global RandomState seeded,
# so the golden also pins the ORDER in which the body consumes it
ssvm = extract('trend_generators/scan_speed_varations.py', ['SSVModulatedSine'])
for tag, seed, amp, per, blip, rt, rate in (
        ('a', 1963, 10, 1.1, 1, [0.278, 7.624, 14.971, 22.317], 0.05),
        ('blip', 7, 5, 0.8, 100, [2.932 * k for k in range(1, 6)], 0.02),
        ('c', 20170410, 10, 1.1, 0, [0.278, 7.624, 14.971, 22.317], 0.1)):
    np.random.seed(seed)
    d_, ri_ = ssvm['SSVModulatedSine'](amp, per, blip).get_subsample_exposure_times(
        None, None, np.array(rt) * U.s, rate * U.s)
    out['ssvm_%s_args' % tag] = np.array([seed, amp, per, blip, rate], dtype=float)
    out['ssvm_%s_rt' % tag] = np.array(rt)
    out['ssvm_%s_dur_ms' % tag] = np.asarray(U.value_in(d_, U.ms), dtype=float)
    out['ssvm_%s_ri' % tag] = np.array(ri_)
    out['ssvm_%s_next_random' % tag] = np.float64(np.random.random())     # stream position afterwards

vt = extract('trend_generators/visit_trends.py', ['BaseVisitTrend', 'HookAndLongTermRamp',
                                                  'gen_orbit_start_times_per_exp'], ns={'abc': __import__('abc')})
t = np.sort(rng.uniform(2456196.1, 2456196.5, 40))
starts = [0, 13, 27]
out['trend_t'] = t
out['trend_t0'] = vt['gen_orbit_start_times_per_exp'](t, starts)
out['trend_factors'] = vt['HookAndLongTermRamp'].ramp_model(t, out['trend_t0'], 0.005, 0.0011, 400, 2456196.28836)

# ---- exposure_generator.py: sample timing ------------------------------------------------------
eg = extract('exposure_generator.py', ['ExposureGenerator'],
             ns={'time': None, 'warnings': None, 'fits': None, 'detector': None, 'exposure': None,
                 'filters': None, 'pyp': None, 'tools': None, 'cosmic_rays': None,
                 'scan_speed_varations': None})
for tag, rt, rate in (('c1', [0.278, 7.624, 14.971, 22.317], 10.0),
                      ('rapid1024', [2.932 * k for k in range(1, 15)], 10.0),
                      ('staring', [0.278, 7.624, 14.971, 22.317], 365.25 * 86400e3)):
    me = types.SimpleNamespace(read_times=np.array(rt) * U.s)
    starts_, mids, durs, ri = eg['ExposureGenerator']._gen_scanning_sample_times(me, rate * U.ms)
    out['times_%s_mid' % tag] = np.asarray(U.value_in(mids, U.ms), dtype=float)
    out['times_%s_dur' % tag] = np.asarray(U.value_in(durs, U.ms), dtype=float)
    out['times_%s_ri' % tag] = np.array(ri)
    out['times_%s_rt' % tag] = np.array(rt)
    yrefs = eg['ExposureGenerator']._gen_sample_yref(me, 457.4, mids, (7.4325 * U.pixel / U.s).to(U.pixel / U.ms))
    out['times_%s_yref' % tag] = np.asarray(yrefs, dtype=float)

# ---- exposure_generator.py: direct_image (:83-144) ----------------------------------------------
class _RecExposure(object):
    def __init__(self, det_, filt, planet, exp_info):
        self.filt, self.exp_info, self.reads = filt, dict(exp_info), []

    def add_read(self, data, read_info=None):
        self.reads.append((np.array(data, dtype=float), read_info))


eg['exposure'] = types.SimpleNamespace(Exposure=_RecExposure)
eg['filters'] = types.SimpleNamespace(F140W=lambda: types.SimpleNamespace(name='F140W'))
eg['ExposureGenerator'].direct_image.__globals__.update(exposure=eg['exposure'], filters=eg['filters'])
for tag, sub, xr_, yr_ in (('256', 256, 404.497, 457.427), ('512', 512, 390.25, 610.75), ('64', 64, 506.1, 498.3)):
    me = types.SimpleNamespace(
        exp_info={}, planet=None, SUBARRAY=sub,
        detector=types.SimpleNamespace(gen_pixel_array=lambda s_, light_sensitive=True, _d=d['WFC3_IR']:
                                       _d.gen_pixel_array(None, s_, light_sensitive)))
    ex = eg['ExposureGenerator'].direct_image(me, xr_, yr_)
    assert len(ex.reads) == 2 and ex.filt.name == 'F140W'
    out['direct_%s_args' % tag] = np.array([sub, xr_, yr_])
    out['direct_%s_zero_shape' % tag] = np.array(ex.reads[0][0].shape)
    out['direct_%s_zero_sum' % tag] = np.float64(np.abs(ex.reads[0][0]).sum())
    out['direct_%s_image' % tag] = ex.reads[1][0]
    out['direct_%s_crpix1' % tag] = np.int64(ex.reads[1][1]['CRPIX1'])
    out['direct_%s_info' % tag] = np.array([me.exp_info['NSAMP'], me.exp_info['cosmic_rate'],
                                            me.exp_info['scale_factor']], dtype=float)
    out['direct_%s_obstype' % tag] = np.array(me.exp_info['OBSTYPE'] + '|' + me.exp_info['SAMP-SEQ'])

np.savez_compressed(os.path.join(HERE, 'reference_goldens.npz'), **out)
print('wrote', len(out), 'arrays;', os.path.getsize(os.path.join(HERE, 'reference_goldens.npz')), 'bytes')
