"""The drop-in boundary: struct layouts of include/rt_types.h (ShaderTypes.h:80-145, Raytracing.metal:168-183,
Renderer.swift:547-556) and the exported C-ABI of the three shared libraries. No GPU needed."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from metal4_raytracing_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")

PROBE = r"""
#include <stdio.h>
#include "rt_types.h"
#include "rt_scene.h"
#include "rt_b200.h"
#define S(t) printf("sizeof %s %zu\n", #t, sizeof(t))
#define O(t, f) printf("offsetof %s.%s %zu\n", #t, #f, offsetof(t, f))
int main(void) {
  S(rt_float3); S(rt_camera); S(rt_light); S(rt_uniforms); S(rt_material); S(rt_instance_descriptor);
  S(rt_texture2d); S(rt_resource); S(rt_image); S(rt_triangle_geometry); S(rt_scene_submesh); S(rt_scene_mesh);
  S(rt_scene_texture); S(rt_scene_instance); S(rt_scene_desc); S(rt_trace_options); S(rt_as_info); S(rt_environment);
  O(rt_environment, intensity); O(rt_environment, flags); O(rt_environment, cdfDev);
  O(rt_camera, right); O(rt_camera, up); O(rt_camera, forward);
  O(rt_light, position); O(rt_light, color); O(rt_light, forward); O(rt_light, right); O(rt_light, up);
  O(rt_light, coneAngle); O(rt_light, direction);
  O(rt_uniforms, width); O(rt_uniforms, height); O(rt_uniforms, blocksWide); O(rt_uniforms, frameIndex);
  O(rt_uniforms, lightCount); O(rt_uniforms, samplesPerPixel); O(rt_uniforms, maxBounces); O(rt_uniforms, camera);
  O(rt_uniforms, previousCamera); O(rt_uniforms, debugTextureMode); O(rt_uniforms, accumulationWeight);
  O(rt_uniforms, enableDenoiseGBuffer); O(rt_uniforms, shadingMode); O(rt_uniforms, enableMotionAdaptiveAccumulation);
  O(rt_uniforms, motionAccumulationMinWeight); O(rt_uniforms, motionAccumulationLowThresholdPixels);
  O(rt_uniforms, motionAccumulationHighThresholdPixels); O(rt_uniforms, enableMotionAdaptiveSampling);
  O(rt_uniforms, motionSamplingMaxExtraSamples); O(rt_uniforms, motionSamplingLowThresholdPixels);
  O(rt_uniforms, motionSamplingHighThresholdPixels);
  O(rt_material, baseColor); O(rt_material, specular); O(rt_material, emission); O(rt_material, specularExponent);
  O(rt_material, refractionIndex); O(rt_material, opacity); O(rt_material, textureFlags);
  O(rt_instance_descriptor, options); O(rt_instance_descriptor, mask);
  O(rt_instance_descriptor, intersectionFunctionTableOffset); O(rt_instance_descriptor, userID);
  O(rt_instance_descriptor, accelerationStructureID);
  O(rt_resource, positions); O(rt_resource, previousPositions); O(rt_resource, normals); O(rt_resource, indices);
  O(rt_resource, material); O(rt_resource, uvs); O(rt_resource, baseColorMap); O(rt_resource, normalMap);
  O(rt_resource, roughnessMap); O(rt_resource, metallicMap); O(rt_resource, aoMap); O(rt_resource, opacityMap);
  O(rt_resource, emissionMap);
  return 0;
}
"""

# values the reference's layouts fix (verified against <simd/simd.h> alignment rules)
REFERENCE_LAYOUT = {
    "sizeof rt_camera": 64, "sizeof rt_light": 128, "sizeof rt_uniforms": 208, "sizeof rt_material": 64,
    "sizeof rt_instance_descriptor": 72, "sizeof rt_resource": 104,
    "offsetof rt_light.position": 16, "offsetof rt_light.color": 32, "offsetof rt_light.forward": 48,
    "offsetof rt_light.right": 64, "offsetof rt_light.up": 80, "offsetof rt_light.coneAngle": 96,
    "offsetof rt_light.direction": 112,
    "offsetof rt_uniforms.frameIndex": 12, "offsetof rt_uniforms.lightCount": 16,
    "offsetof rt_uniforms.samplesPerPixel": 20, "offsetof rt_uniforms.maxBounces": 24, "offsetof rt_uniforms.camera": 32,
    "offsetof rt_uniforms.previousCamera": 96, "offsetof rt_uniforms.debugTextureMode": 160,
    "offsetof rt_uniforms.accumulationWeight": 164, "offsetof rt_uniforms.enableDenoiseGBuffer": 168,
    "offsetof rt_uniforms.shadingMode": 172, "offsetof rt_uniforms.enableMotionAdaptiveAccumulation": 176,
    "offsetof rt_uniforms.motionAccumulationMinWeight": 180, "offsetof rt_uniforms.motionAccumulationLowThresholdPixels": 184,
    "offsetof rt_uniforms.motionAccumulationHighThresholdPixels": 188, "offsetof rt_uniforms.enableMotionAdaptiveSampling": 192,
    "offsetof rt_uniforms.motionSamplingMaxExtraSamples": 196, "offsetof rt_uniforms.motionSamplingLowThresholdPixels": 200,
    "offsetof rt_uniforms.motionSamplingHighThresholdPixels": 204,
    "offsetof rt_material.specular": 16, "offsetof rt_material.emission": 32, "offsetof rt_material.specularExponent": 48,
    "offsetof rt_material.refractionIndex": 52, "offsetof rt_material.opacity": 56, "offsetof rt_material.textureFlags": 60,
    "offsetof rt_instance_descriptor.options": 48, "offsetof rt_instance_descriptor.mask": 52,
    "offsetof rt_instance_descriptor.userID": 60, "offsetof rt_instance_descriptor.accelerationStructureID": 64,
    "offsetof rt_resource.indices": 24, "offsetof rt_resource.material": 32, "offsetof rt_resource.uvs": 40,
    "offsetof rt_resource.baseColorMap": 48, "offsetof rt_resource.emissionMap": 96,
}

CTYPES = {
    "rt_float3": A.Float3, "rt_camera": A.Camera, "rt_light": A.Light, "rt_uniforms": A.Uniforms,
    "rt_material": A.Material, "rt_instance_descriptor": A.InstanceDescriptor, "rt_texture2d": A.Texture2D,
    "rt_resource": A.Resource, "rt_image": A.Image, "rt_triangle_geometry": A.TriangleGeometry,
    "rt_scene_submesh": A.SceneSubmesh, "rt_scene_mesh": A.SceneMesh, "rt_scene_texture": A.SceneTexture,
    "rt_scene_instance": A.SceneInstance, "rt_scene_desc": A.SceneDesc,
}


@pytest.fixture(scope="module")
def c_layout():
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "probe.c")
        with open(src, "w") as f:
            f.write(PROBE)
        exe = os.path.join(d, "probe")
        subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-I", INC, src, "-o", exe])
        out = subprocess.check_output([exe], text=True)
    layout = {}
    for line in out.splitlines():
        kind, name, val = line.split()
        layout[f"{kind} {name}"] = int(val)
    return layout


def test_header_matches_reference_layout(c_layout):
    for key, val in REFERENCE_LAYOUT.items():
        assert c_layout[key] == val, key


def test_ctypes_mirror_matches_header(c_layout):
    from metal4_raytracing_b200 import device
    for cname, ctype in CTYPES.items():
        assert C.sizeof(ctype) == c_layout[f"sizeof {cname}"], cname
    assert C.sizeof(device.TraceOptions) == c_layout["sizeof rt_trace_options"]
    assert C.sizeof(device.AsInfo) == c_layout["sizeof rt_as_info"]
    assert C.sizeof(device.Environment) == c_layout["sizeof rt_environment"] == 32
    for field in ("intensity", "flags", "cdfDev"):
        assert getattr(device.Environment, field).offset == c_layout[f"offsetof rt_environment.{field}"], field
    for key, val in c_layout.items():
        if not key.startswith("offsetof") or key.startswith("offsetof rt_environment"):
            continue
        cname, field = key.split()[1].split(".")
        assert getattr(CTYPES[cname], field).offset == val, key


def test_binding_indices_match_reference():
    # ShaderTypes.h:35-70
    assert (A.BUFFER_UNIFORMS, A.BUFFER_RESOURCES, A.BUFFER_LIGHTS, A.BUFFER_ACCELERATION_STRUCTURE,
            A.BUFFER_INSTANCE_DESCRIPTORS) == (0, 5, 6, 8, 9)
    assert [A.BUFFER_REST_POSITIONS, A.BUFFER_REST_NORMALS, A.BUFFER_JOINT_INDICES, A.BUFFER_JOINT_WEIGHTS,
            A.BUFFER_JOINT_MATRICES, A.BUFFER_SKINNED_POSITIONS, A.BUFFER_SKINNED_NORMALS] == list(range(10, 17))
    assert A.BUFFER_PREVIOUS_INSTANCE_DESCRIPTORS == 17 and A.BUFFER_COUNT == 18
    assert [A.TEXTURE_ACCUMULATION, A.TEXTURE_PREVIOUS_ACCUMULATION, A.TEXTURE_RANDOM, A.TEXTURE_DEPTH,
            A.TEXTURE_MOTION, A.TEXTURE_DIFFUSE_ALBEDO, A.TEXTURE_SPECULAR_ALBEDO, A.TEXTURE_NORMAL,
            A.TEXTURE_ROUGHNESS] == list(range(9))
    assert (A.LIGHT_SUN, A.LIGHT_SPOT, A.LIGHT_POINT, A.LIGHT_AREA) == (1, 2, 3, 4)


def _declared(header, prefix):
    text = open(os.path.join(INC, header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(%s_[a-z0-9_]+)\s*\(" % prefix, text)))


@pytest.mark.parametrize("header,prefix,libname", [
    ("rt_b200.h", "rt", "librt_b200.so"), ("rt_renderer.h", "rtr", "librt_b200.so"),
    ("rt_scene.h", "rts", "librt_scene.so")])
def test_library_exports_every_declared_symbol(header, prefix, libname):
    """dlopen works without a GPU and every function the header declares resolves."""
    names = _declared(header, prefix)
    assert len(names) >= 10
    lib = C.CDLL(os.path.join(ROOT, "metal4_raytracing_b200", "lib", libname))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_export_list_is_complete():
    from metal4_raytracing_b200 import device
    declared = set(_declared("rt_b200.h", "rt") + _declared("rt_renderer.h", "rtr"))
    assert declared - set(device.EXPORTS) <= {"rt_pack_tiles", "rt_unpack_tiles", "rt_ipc_export", "rt_ipc_import",
                                               "rt_ipc_close", "rt_selftest_child_boxes"}
    assert set(device.EXPORTS) <= declared


def test_no_gpu_means_loud_failure():
    """Without a CUDA device the product refuses to run (no CPU fallback)."""
    from metal4_raytracing_b200 import device
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(device.RtError):
        device.Context(0)


def test_product_does_not_import_oracle():
    """Nothing under the package may reference oracle/ (only tests, smoke() and bench.py's CPU legs do)."""
    pkg = os.path.join(ROOT, "metal4_raytracing_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                code = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith(("//", "#", "*", "/*")))
                assert "import oracle" not in code and "liboracle" not in code and '"oracle' not in code, fn
                assert "../oracle" not in code and "oracle/" not in code, fn


def test_cpp_host_program_fails_loudly_without_gpu():
    """The C++ example client (examples/rt_render.cpp) links both libraries through the C headers; without a CUDA
    device it exits non-zero with rt_create's message instead of falling back to anything."""
    import subprocess
    import torch
    from metal4_raytracing_b200 import device
    exe = os.path.join(os.path.dirname(device.LIB_PATH), "rt_render")
    assert os.path.isfile(exe)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([exe, "K3small", "32", "32", "1", "1", "1", "/tmp/_rt_render_should_not_exist.png"],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode != 0 and "no CPU path" in out.stderr
    assert not os.path.exists("/tmp/_rt_render_should_not_exist.png")
