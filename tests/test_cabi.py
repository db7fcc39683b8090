"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/nf_b200.h
declares, size queries work without a device, and the host-side mirror of the reference interface has
the reference's state-dict layout.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

import neurofluid_b200 as nb
from neurofluid_b200 import _lib, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "nf_b200.h")).read()
    return sorted(set(re.findall(r"NF_API\s+[\w\s\*]+?\b(nf_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 15
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_lib.SIGNATURES) == names          # the ctypes binding covers the header exactly


def test_size_queries_and_version_without_device():
    L = _lib.lib()
    assert L.nf_version() == 100
    assert L.nf_render_packed_weights_bytes() == 154 * 8192 + 20 * 4096 + 3208 * 4
    assert L.nf_grid_workspace_bytes(1000) > 3 * 64 ** 3 * 4
    assert L.nf_render_workspace_bytes(1024, 64, 128) > 1024 * 192 * (64 + 16)
    assert L.nf_render_workspace_bytes(0, 64, 128) == 0


STRUCTS = {"nf_render_args": "RenderArgs", "nf_transition_args": "TransitionArgs", "nf_render_ws_view": "RenderWsView",
           "nf_cconv_args": "CConvArgs", "nf_render_bwd_args": "RenderBwdArgs",
           "nf_transition_bwd_args": "TransitionBwdArgs"}


def test_struct_layout_matches_header(tmp_path):
    """Every ctypes mirror has the size and field offsets the C compiler gives the header's struct (gcc compiles a
    probe that prints offsetof() for every field)."""
    import subprocess
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "nf_b200.h"', 'int main(void) {']
    for cname, pyname in STRUCTS.items():
        cls = getattr(_lib, pyname)
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['return 0; }']
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, pyname in STRUCTS.items():
        cls = getattr(_lib, pyname)
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


def test_rendernet_state_dict_layout_and_errors():
    net = nb.RenderNet(scenes.render_cfg(), scenes.NEAR, scenes.FAR)
    sd = scenes.init_render_state(0)
    assert sorted(net.state_dict().keys()) == sorted(sd.keys())
    net.load_state_dict(sd, strict=True)
    assert net.nerf_coarse.xyz_encoding_5[0].weight.shape == (256, 454)
    assert net.nerf_fine.dir_encoding[0].weight.shape == (128, 310)
    assert nb.Renderer is nb.RenderNet
    cw = torch.from_numpy(scenes.CAMERA_C2W)
    assert torch.equal(net.set_ro(cw), cw[:, 3])
    with pytest.raises(_lib.NFError):                 # CPU tensors never fall back to a CPU path
        net(torch.zeros(10, 3), cw[:, 3], torch.zeros(4, 6), 1.0, cw)
    assert nb.RenderNet(scenes.render_cfg(exclude_ray=False), scenes.NEAR, scenes.FAR).include_ray
    # encoding ablations narrow the networks (models/renderer.py:25-42): state-dict shapes follow the reference's
    abl = nb.RenderNet(scenes.render_cfg(var=False, smoothed_dir=False), scenes.NEAR, scenes.FAR)
    assert abl.nerf_fine.xyz_encoding_1[0].weight.shape == (256, 135) and abl.nerf_fine.dir_encoding[0].weight.shape == (128, 283)
    assert abl.nerf_fine.xyz_encoding_5[0].weight.shape == (256, 135 + 256)


def test_particlenet_state_dict_layout():
    net = nb.ParticleNet(gravity=(0.0, 0.0, -9.81))
    sd = scenes.init_particle_state(0)
    assert sorted(net.state_dict().keys()) == sorted(sd.keys())
    net.load_state_dict(sd, strict=True)
    assert net.conv1.kernel.shape == (4, 4, 4, 96, 64) and net.conv3.kernel.shape == (4, 4, 4, 64, 3)
    assert net.dense1.weight.shape == (64, 96) and tuple(net.gravity.tolist()) == (0.0, 0.0, pytest.approx(-9.81))
    assert float(net.filter_extent) == pytest.approx(0.225)
    L = _lib.lib()
    assert L.nf_transition_num_phases() == 5
    assert L.nf_transition_workspace_bytes(1000, 500) > 2 * 1000 * 128 * 48
    assert ctypes.sizeof(_lib.TransitionArgs) % 8 == 0


def test_comm_entry_points_are_no_ops_for_one_rank():
    """Without nf_comm_init the library is a world of one: the collective entry points return at once (no CUDA call)."""
    L = _lib.lib()
    r, w = ctypes.c_int(-1), ctypes.c_int(-1)
    assert L.nf_comm_info(ctypes.byref(r), ctypes.byref(w)) == 0 and (r.value, w.value) == (0, 1)
    assert L.nf_comm_register_buffer(None, 0) == 0
    assert L.nf_allgather_rows(None, 0, None) == 0
    n = ctypes.c_uint(7)
    assert L.nf_comm_exchange_timeouts(ctypes.byref(n)) == 0 and n.value == 0
