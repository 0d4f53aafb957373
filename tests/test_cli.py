"""The lc_cutout CLI / processLC layer (cosmo-cutout_b200/host) against the reference binary built from
/root/reference (oracle/_ref/lc_cutout_ref, travels prebuilt to the GPU box): same command line, same input
directory, byte-identical output directories."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

import harness as H
from __graft_entry__ import REPO, load_package

cc = load_package()
REF_BIN = os.path.join(REPO, "oracle", "_ref", "lc_cutout_ref")
needs_ref_bin = pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/lc_cutout_ref not built")


def _run(binary, args, env=None, check=True):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([binary] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=e,
                       text=True, timeout=600)
    if check:
        assert p.returncode == 0, p.stdout[-3000:]
    return p


def _same_tree(a, b):
    """every file under a exists under b with identical bytes, and vice versa"""
    fa = sorted(os.path.relpath(os.path.join(r, f), a) for r, _, fs in os.walk(a) for f in fs)
    fb = sorted(os.path.relpath(os.path.join(r, f), b) for r, _, fs in os.walk(b) for f in fs)
    assert fa == fb, (fa, fb)
    for f in fa:
        assert filecmp.cmp(os.path.join(a, f), os.path.join(b, f), shallow=False), f
    return fa


def _halo_file(path, rows):
    with open(path, "w") as f:
        for (tag, pos) in rows:
            f.write(f"{tag} 0.21 421 3.1e14 1.25 4.5 0.3 {pos[0]} {pos[1]} {pos[2]}\n")


@pytest.fixture()
def lc_dirs(tmp_path):
    inp = tmp_path / "lc"
    inp.mkdir()
    for step, seed in ((487, 1), (475, 2), (499, 3)):
        cols = H.concat_cols(H.shell_fixture(150_000, seed=seed), H.shell_fixture(50_000, seed=seed + 10, octant=False))
        H.write_flat_step(str(inp), step, cols)
    return str(inp)


# ------------------------------------------------------------------------------------------------ CPU
@needs_ref_bin
def test_props_only_equals_reference(lc_dirs, tmp_path):
    """--propsOnly needs no GPU: properties.csv (processLC.cpp:740-769) must be the reference's bytes"""
    for sub, extra in (("a", ["--halo", 150, 20, 12, "--boxLength", 10]),
                       ("b", ["-h", 30.5, 170.25, 100, "-b", 30])):
        o1, o2 = tmp_path / ("ours_" + sub), tmp_path / ("ref_" + sub)
        o1.mkdir(); o2.mkdir()
        _run(cc.CLI_PATH, [lc_dirs, o1, 487, 470] + extra + ["--propsOnly"])
        _run(REF_BIN, [lc_dirs, o2, 487, 470] + extra + ["--propsOnly"])
        assert _same_tree(str(o1), str(o2)) == ["properties.csv"]
    hf = tmp_path / "halos.txt"
    _halo_file(str(hf), [("1001", (150, 20, 12)), ("1002", (30, 170, 100)), ("77", (100, 100, 100))])
    o1, o2 = tmp_path / "ours_f", tmp_path / "ref_f"
    o1.mkdir(); o2.mkdir()
    _run(cc.CLI_PATH, [lc_dirs, o1, 0.0, 0.06, "-f", hf, "-b", 15, "--propsOnly"])
    _run(REF_BIN, [lc_dirs, o2, 0.0, 0.06, "-f", hf, "-b", 15, "--propsOnly"])
    assert len(_same_tree(str(o1), str(o2))) == 3


def test_cli_argument_errors(lc_dirs, tmp_path):
    out = tmp_path / "o"
    out.mkdir()
    for bad in (["--halo", 1, 2, 3], ["--theta", 87, 2], ["--theta", 87, 2, "--phi", 2, 2, "--halo", 1, 2, 3, "-b", 5], [],
                ["-m", "sod", "--halo", 1, 2, 3, "-b", 5]):
        p = _run(cc.CLI_PATH, [lc_dirs, out, 487, 470] + bad, check=False)
        assert p.returncode != 0, bad
    p = _run(cc.CLI_PATH, [lc_dirs, tmp_path / "missing", 487, 470, "--theta", 87, 2, "--phi", 2, 2], check=False)
    assert p.returncode != 0 and "Cannot access specified output directory" in p.stdout


def test_cli_without_gpu_fails_loudly(lc_dirs, tmp_path):
    if cc.lib().cc_device_count() > 0:
        pytest.skip("a GPU is present")
    out = tmp_path / "o"
    out.mkdir()
    p = _run(cc.CLI_PATH, [lc_dirs, out, 487, 470, "--theta", 87, 2, "--phi", 2, 2], check=False)
    assert p.returncode != 0 and "no CUDA device" in p.stdout


@needs_ref_bin
def test_step_selection_equals_reference(lc_dirs, tmp_path):
    """the step list printed by main (zToStep/stepToZ/getLCSteps) matches, for redshift and snapshot bounds"""
    o1, o2 = tmp_path / "o1", tmp_path / "o2"
    o1.mkdir(); o2.mkdir()
    for bounds in ((487, 470), (0.0, 0.06), (499, 480), (0.02, 0.03)):
        a = _run(cc.CLI_PATH, [lc_dirs, o1, *bounds, "--halo", 150, 20, 12, "-b", 10, "--propsOnly"]).stdout
        b = _run(REF_BIN, [lc_dirs, o2, *bounds, "--halo", 150, 20, 12, "-b", 10, "--propsOnly"]).stdout

        def grab(txt):
            lines = txt.splitlines()
            i = next(k for k, l in enumerate(lines) if l.startswith("MAX STEP"))
            return lines[i:i + 4]
        assert grab(a) == grab(b), bounds


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@needs_ref_bin
@pytest.mark.parametrize("ngpus", [1, 2])
def test_cli_const_equals_reference(lc_dirs, tmp_path, ngpus):
    if ngpus > cc.lib().cc_device_count():
        pytest.skip(f"{ngpus} GPUs needed")
    o1, o2 = tmp_path / "ours", tmp_path / "ref"
    o1.mkdir(); o2.mkdir()
    args = [487, 470, "--theta", 80, 8, "--phi", 40, 10]
    _run(cc.CLI_PATH, [lc_dirs, o1] + args, env={"CC_NGPUS": str(ngpus)})
    _run(REF_BIN, [lc_dirs, o2] + args)
    files = _same_tree(str(o1), str(o2))
    assert len(files) == 24 and os.path.getsize(os.path.join(str(o1), "lcCutout487", "id.487.bin")) > 8 * 100
    # a second run into the now non-empty directory must refuse (processLC.cpp:174-177)
    p = _run(cc.CLI_PATH, [lc_dirs, o1] + args, check=False)
    assert p.returncode != 0 and "is non-empty" in p.stdout


@pytest.mark.gpu
@needs_ref_bin
@pytest.mark.parametrize("pos_only", [False, True])
@pytest.mark.parametrize("ngpus", [1, 2])
def test_cli_halo_equals_reference(lc_dirs, tmp_path, pos_only, ngpus):
    if ngpus > cc.lib().cc_device_count():
        pytest.skip(f"{ngpus} GPUs needed")
    gpu_env = {"CC_NGPUS": str(ngpus)}
    hf = tmp_path / "halos.txt"
    _halo_file(str(hf), [("1001", (150, 20, 12)), ("1002", (30, 170, 100)), ("77", (100, 100, 100)),
                         ("5", (10, 10, 250)), ("6", (120, -60, 40))])
    o1, o2 = tmp_path / "ours", tmp_path / "ref"
    o1.mkdir(); o2.mkdir()
    args = [0.0, 0.06, "-f", hf, "-b", 120] + (["--posOnly"] if pos_only else [])
    a = _run(cc.CLI_PATH, [lc_dirs, o1] + args + ["--timeit"], env=gpu_env)
    _run(REF_BIN, [lc_dirs, o2] + args)
    files = _same_tree(str(o1), str(o2))
    assert len(files) == 5 * (1 + 2 * (7 if pos_only else 12))
    assert "cutout_times = np.array([" in a.stdout
    sizes = [os.path.getsize(os.path.join(str(o1), "halo_1001", "lcCutout487", "id.487.bin"))]
    assert sizes[0] > 0
    # re-run without --overwrite: every halo is skipped and nothing changes; with --overwrite: identical again
    b = _run(cc.CLI_PATH, [lc_dirs, o1] + args, env=gpu_env)
    assert "skipping halo" in b.stdout
    _same_tree(str(o1), str(o2))
    _run(cc.CLI_PATH, [lc_dirs, o1] + args + ["--overwrite"], env=gpu_env)
    _same_tree(str(o1), str(o2))


@pytest.mark.gpu
@needs_ref_bin
def test_cli_single_halo_equals_reference(lc_dirs, tmp_path):
    o1, o2 = tmp_path / "ours", tmp_path / "ref"
    o1.mkdir(); o2.mkdir()
    args = [487, 470, "--halo", 150, 20, 12, "--boxLength", 90, "-v"]
    _run(cc.CLI_PATH, [lc_dirs, o1] + args)
    _run(REF_BIN, [lc_dirs, o2] + args)
    assert len(_same_tree(str(o1), str(o2))) == 1 + 2 * 12


@pytest.mark.gpu
@needs_ref_bin
def test_cli_reads_genericio_format_input(tmp_path):
    """the same steps once as flat columns (read by the reference binary through its shimmed reader) and once in the
    GenericIO on-disk format -- a single file, a partitioned big-endian one, one with block headers -- give the same
    output trees, for both use cases"""
    flat, gio = tmp_path / "flat", tmp_path / "gio"
    flat.mkdir(); gio.mkdir()
    variants = {499: dict(nranks=4), 475: dict(nranks=6, partitions=3, big_endian=True), 487: dict(nranks=3, block_headers=True)}
    for step, seed in ((487, 1), (475, 2), (499, 3)):
        cols = H.concat_cols(H.shell_fixture(120_000, seed=seed), H.shell_fixture(40_000, seed=seed + 10, octant=False))
        H.write_flat_step(str(flat), step, cols)
        H.write_gio_step(str(gio), step, cols, seed=seed, **variants[step])
    for name, args in (("const", [499, 470, "--theta", 80, 8, "--phi", 40, 10]),
                       ("halo", [499, 470, "--halo", 150, 20, 12, "--boxLength", 90])):
        o1, o2 = tmp_path / ("ours_" + name), tmp_path / ("ref_" + name)
        o1.mkdir(); o2.mkdir()
        _run(cc.CLI_PATH, [str(gio), o1] + args)
        _run(REF_BIN, [str(flat), o2] + args)
        files = _same_tree(str(o1), str(o2))
        assert len(files) >= 2 * 12 and any("475" in f for f in files) and any("487" in f for f in files)


@pytest.mark.gpu
@needs_ref_bin
@pytest.mark.parametrize("passes", ["shared", "separate"])
@pytest.mark.parametrize("ngpus", [1, 2])
def test_cli_several_box_lengths_keep_steps_resident(lc_dirs, tmp_path, ngpus, passes):
    """extension over the reference CLI: --boxLength with several values (the reference's API has one boxLength per
    call, processLC.h:46; BASELINE config 5).  Each value's tree equals a separate reference run with that value.
    Default ("shared"): all (halo, box length) pairs share ONE device pass per lightcone step (processLC_boxes) --
    every step is read and uploaded once.  CC_BOX_PASSES=separate: one processLC call per value; after the first
    call the steps are found resident in HBM, the later calls copy ZERO bytes host->device and read no input file."""
    if ngpus > cc.lib().cc_device_count():
        pytest.skip(f"{ngpus} GPUs needed")
    import re
    hf = tmp_path / "halos.txt"
    _halo_file(str(hf), [("1001", (150, 20, 12)), ("1002", (30, 170, 100)), ("77", (100, 100, 100)),
                         ("5", (10, 10, 250)), ("6", (120, -60, 40)), ("7", (200, 120, 40)), ("8", (60, 60, 20)),
                         ("9", (250, 30, 160)), ("10", (90, 200, 30))])
    boxes = ["30", "120", "75.5"]
    ours = tmp_path / "ours"
    ours.mkdir()
    env = {"CC_NGPUS": str(ngpus)}
    if passes == "separate":
        env["CC_BOX_PASSES"] = "separate"
    a = _run(cc.CLI_PATH, [lc_dirs, ours, 0.0, 0.06, "-f", hf, "-b"] + boxes, env=env)
    for b in boxes:
        ref = tmp_path / f"ref_{b}"
        ref.mkdir()
        _run(REF_BIN, [lc_dirs, ref, 0.0, 0.06, "-f", hf, "-b", b])
        files = _same_tree(str(ours / f"boxLength_{b}"), str(ref))
        assert len(files) == 9 * (1 + 2 * 12)
    moved = [int(m) for m in re.findall(r"step columns copied host->device in this call: (\d+) bytes", a.stdout)]
    found = re.findall(r"resident steps: (\d+) found in HBM, (\d+) read and uploaded", a.stdout)
    if passes == "separate":
        assert len(moved) == 3 and moved[0] == 2 * 200_000 * 44 and moved[1] == 0 and moved[2] == 0, moved
        assert found == [("0", "2"), ("2", "0"), ("2", "0")], found
    else:
        # one call: x,y,z of both steps streamed once for the 27 cutouts (the other columns are picked on the host)
        assert len(moved) == 1 and moved[0] == 2 * 200_000 * 12, moved
        return
    # use case 1 shares the cache machinery (CC_RESIDENT_STEPS=1): same bytes as the streamed default
    o1, o2 = tmp_path / "c1", tmp_path / "c2"
    o1.mkdir(); o2.mkdir()
    args = [0.0, 0.06, "--theta", 87, 2, "--phi", 2, 2]
    _run(cc.CLI_PATH, [lc_dirs, o1] + args, env={"CC_NGPUS": str(ngpus), "CC_RESIDENT_STEPS": "1"})
    _run(REF_BIN, [lc_dirs, o2] + args)
    _same_tree(str(o1), str(o2))
