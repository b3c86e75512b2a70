"""ctypes mirrors of include/rt_types.h and include/rt_scene.h.

The layouts are the reference's ShaderTypes.h structs (MetalRaytracing/ShaderTypes.h:80-145), the Apple
instance descriptor it fills (MetalRaytracing/Renderer.swift:547-556) and the Resource argument-buffer row
(MetalRaytracing/Raytracing.metal:168-183). tests/test_abi.py checks every size/offset against the C header.
"""
import ctypes as C

BUFFER_UNIFORMS = 0
BUFFER_RESOURCES = 5
BUFFER_LIGHTS = 6
BUFFER_ACCELERATION_STRUCTURE = 8
BUFFER_INSTANCE_DESCRIPTORS = 9
BUFFER_REST_POSITIONS = 10
BUFFER_REST_NORMALS = 11
BUFFER_JOINT_INDICES = 12
BUFFER_JOINT_WEIGHTS = 13
BUFFER_JOINT_MATRICES = 14
BUFFER_SKINNED_POSITIONS = 15
BUFFER_SKINNED_NORMALS = 16
BUFFER_PREVIOUS_INSTANCE_DESCRIPTORS = 17
BUFFER_COUNT = 18

TEXTURE_ACCUMULATION = 0
TEXTURE_PREVIOUS_ACCUMULATION = 1
TEXTURE_RANDOM = 2
TEXTURE_DEPTH = 3
TEXTURE_MOTION = 4
TEXTURE_DIFFUSE_ALBEDO = 5
TEXTURE_SPECULAR_ALBEDO = 6
TEXTURE_NORMAL = 7
TEXTURE_ROUGHNESS = 8
TEXTURE_COUNT = 9

LIGHT_SUN, LIGHT_SPOT, LIGHT_POINT, LIGHT_AREA = 1, 2, 3, 4
SHADING_PBR, SHADING_LEGACY = 0, 1
# DebugTextureMode (ShaderTypes.h:159-168)
DEBUG_NONE, DEBUG_BASECOLOR, DEBUG_NORMAL, DEBUG_ROUGHNESS, DEBUG_METALLIC, DEBUG_AO, DEBUG_EMISSION, DEBUG_MOTION = range(8)

FORMAT_NONE = 0
FORMAT_R32_UINT = 1
FORMAT_R32_FLOAT = 2
FORMAT_RG16_FLOAT = 3
FORMAT_RGBA16_FLOAT = 4
FORMAT_R16_FLOAT = 5
FORMAT_RG32_FLOAT = 6
FORMAT_RGBA32_FLOAT = 7

AS_FLAG_COMPACT = 1
AS_FLAG_REFITTABLE = 2
INTERSECT_ANY = 1  # rt_intersect flags

SLOT_BASECOLOR, SLOT_NORMAL, SLOT_ROUGHNESS, SLOT_METALLIC, SLOT_AO, SLOT_OPACITY, SLOT_EMISSION = range(7)


class Float3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("_pad", C.c_float)]

    def set(self, x, y, z):
        self.x, self.y, self.z, self._pad = float(x), float(y), float(z), 0.0

    def tuple(self):
        return (self.x, self.y, self.z)


class Camera(C.Structure):
    _fields_ = [("position", Float3), ("right", Float3), ("up", Float3), ("forward", Float3)]


class Light(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("_pad0", C.c_int32 * 3),
        ("position", Float3),
        ("color", Float3),
        ("forward", Float3),
        ("right", Float3),
        ("up", Float3),
        ("coneAngle", C.c_float),
        ("_pad1", C.c_float * 3),
        ("direction", Float3),
    ]


class Uniforms(C.Structure):
    _fields_ = [
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("blocksWide", C.c_int32),
        ("frameIndex", C.c_uint32),
        ("lightCount", C.c_int32),
        ("samplesPerPixel", C.c_int32),
        ("maxBounces", C.c_int32),
        ("_pad0", C.c_int32),
        ("camera", Camera),
        ("previousCamera", Camera),
        ("debugTextureMode", C.c_int32),
        ("accumulationWeight", C.c_float),
        ("enableDenoiseGBuffer", C.c_int32),
        ("shadingMode", C.c_int32),
        ("enableMotionAdaptiveAccumulation", C.c_int32),
        ("motionAccumulationMinWeight", C.c_float),
        ("motionAccumulationLowThresholdPixels", C.c_float),
        ("motionAccumulationHighThresholdPixels", C.c_float),
        ("enableMotionAdaptiveSampling", C.c_int32),
        ("motionSamplingMaxExtraSamples", C.c_int32),
        ("motionSamplingLowThresholdPixels", C.c_float),
        ("motionSamplingHighThresholdPixels", C.c_float),
    ]

    def copy(self):
        u = Uniforms()
        C.memmove(C.byref(u), C.byref(self), C.sizeof(Uniforms))
        return u


class Material(C.Structure):
    _fields_ = [
        ("baseColor", Float3),
        ("specular", Float3),
        ("emission", Float3),
        ("specularExponent", C.c_float),
        ("refractionIndex", C.c_float),
        ("opacity", C.c_float),
        ("textureFlags", C.c_uint32),
    ]


class InstanceDescriptor(C.Structure):
    _fields_ = [
        ("transformationMatrix", (C.c_float * 3) * 4),
        ("options", C.c_uint32),
        ("mask", C.c_uint32),
        ("intersectionFunctionTableOffset", C.c_uint32),
        ("userID", C.c_uint32),
        ("accelerationStructureID", C.c_uint64),
    ]


class Texture2D(C.Structure):
    _fields_ = [("texels", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("srgb", C.c_int32),
                ("_pad", C.c_int32)]


class Resource(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "positions", "previousPositions", "normals", "indices", "material", "uvs", "baseColorMap", "normalMap",
        "roughnessMap", "metallicMap", "aoMap", "opacityMap", "emissionMap")]


class Image(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("format", C.c_int32),
                ("_pad", C.c_int32)]


class TriangleGeometry(C.Structure):
    _fields_ = [("vertexBuffer", C.c_void_p), ("vertexStride", C.c_uint32), ("vertexCount", C.c_uint32),
                ("indexBuffer", C.c_void_p), ("indexStride", C.c_uint32), ("triangleCount", C.c_uint32)]


# ---- rt_scene.h ---------------------------------------------------------------------------------------
class SceneSubmesh(C.Structure):
    _fields_ = [("indices", C.POINTER(C.c_int32)), ("triangleCount", C.c_uint32), ("_pad", C.c_uint32),
                ("material", Material), ("textureIndex", C.c_int32 * 7), ("_pad2", C.c_int32)]


class SceneMesh(C.Structure):
    _fields_ = [("vertexCount", C.c_uint32), ("submeshCount", C.c_uint32),
                ("positions", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)),
                ("uvs", C.POINTER(C.c_float)), ("jointIndices", C.POINTER(C.c_uint16)),
                ("jointWeights", C.POINTER(C.c_float)), ("jointCount", C.c_uint32), ("_pad", C.c_uint32),
                ("jointMatrices", C.POINTER(C.c_float)), ("submeshes", C.POINTER(SceneSubmesh)),
                ("jointParents", C.POINTER(C.c_int32)), ("jointInverseBind", C.POINTER(C.c_float)),
                ("jointLocalTRS", C.POINTER(C.c_float))]


class SceneTexture(C.Structure):
    _fields_ = [("texels", C.POINTER(C.c_uint8)), ("width", C.c_int32), ("height", C.c_int32),
                ("srgb", C.c_int32), ("_pad", C.c_int32)]


class SceneInstance(C.Structure):
    _fields_ = [("meshIndex", C.c_uint32), ("_pad", C.c_uint32), ("transform", C.c_float * 16),
                ("previousTransform", C.c_float * 16)]


class SceneDesc(C.Structure):
    _fields_ = [("meshCount", C.c_uint32), ("textureCount", C.c_uint32), ("instanceCount", C.c_uint32),
                ("lightCount", C.c_uint32), ("maxSubmeshes", C.c_uint32), ("_pad", C.c_uint32),
                ("meshes", C.POINTER(SceneMesh)), ("textures", C.POINTER(SceneTexture)),
                ("instances", C.POINTER(SceneInstance)), ("lights", C.POINTER(Light))]


EXPECTED_SIZES = {
    Float3: 16, Camera: 64, Light: 128, Uniforms: 208, Material: 64, InstanceDescriptor: 72, Texture2D: 24,
    Resource: 104, Image: 24, TriangleGeometry: 32,
}
for _t, _n in EXPECTED_SIZES.items():
    assert C.sizeof(_t) == _n, (_t.__name__, C.sizeof(_t), _n)
