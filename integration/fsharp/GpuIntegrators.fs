// GpuIntegrators.fs — the reference-side binding of libbarnacle_b200.so.
//
// STATUS: source only.  `dotnet` / `fsharpc` are not in the build image, so this file has never been
// compiled or run; the same C-ABI calls, in the same order and with the same struct layouts, are what
// barnacle_b200/_ffi.py + barnacle_b200/scene.py make in every `-m gpu` test (tests/test_abi.py pins the
// struct sizes and offsets quoted below against include/barnacle_b200.h).
//
// Where it goes in the reference tree (LeonKang130/Barnacle):
//   * copy to  Extensions/Integrator/GpuIntegrators.fs
//   * Barnacle.fsproj: <Compile Include="Extensions\Integrator\GpuIntegrators.fs" /> AFTER
//     Extensions\Aggregate\BVH.fs and BEFORE Extensions\Scene\Loader.fs (it needs the material, primitive
//     and camera types, which the project compiles after the CPU integrators)
//   * Extensions/Scene/Loader.fs:185-204 and Extensions/Scene/Render.fs:12-14: the two edits at the end
//     of this file (also in INTEGRATION.md §3)
//   * libbarnacle_b200.so next to the executable (or on LD_LIBRARY_PATH)
//
// What it replaces: IntegratorBase.Render (Base/Integrator.fs:9-11) for the path-tracing, direct, normal
// and PSSMLT integrators.  Everything else of the program (Scene.Load, Scene.Traverse, BVHNode.Build,
// AliasTable, Film.Save) keeps running in .NET and is the source of the flattened arrays.
#nowarn "9" // NativePtr, fixed
#nowarn "51" // address-of
namespace Barnacle.Extensions.Integrator

open System
open System.Collections.Generic
open System.Numerics
open System.Runtime.InteropServices
open Microsoft.FSharp.NativeInterop
open Barnacle.Util
open Barnacle.Base
open Barnacle.Extensions.Primitive
open Barnacle.Extensions.Material
open Barnacle.Extensions.Camera

/// Blittable mirrors of the records in include/barnacle_b200.h (sizes in bytes after each type).
module Native =
    [<Literal>]
    let Lib = "barnacle_b200"

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type BnInstance = // 168
        val mutable primKind: uint32 // 0 mesh, 1 sphere
        val mutable primId: uint32
        val mutable materialId: int // -1: HasMaterial = false
        val mutable lightId: int // -1: HasLight = false
        val mutable objectToWorld: Matrix4x4 // 16 floats row-major, row-vector convention: as is
        val mutable worldToObject: Matrix4x4
        val mutable boundsMin: Vector3
        val mutable boundsMax: Vector3

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type BnMesh = // 32
        val mutable vertexOffset: uint32
        val mutable vertexCount: uint32
        val mutable triOffset: uint32
        val mutable triCount: uint32
        val mutable nodeOffset: uint32
        val mutable nodeCount: uint32
        val mutable aliasOffset: uint32
        val mutable reserved: uint32

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type BnMaterial = // 24
        val mutable kind: uint32 // 0 lambertian, 1 mirror, 2 dielectric, 3 pbr
        val mutable baseColor: Vector3
        val mutable p0: float32 // dielectric: IOR | pbr: Metallic
        val mutable p1: float32 // pbr: Alpha

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type BnLight = // 16
        val mutable emission: Vector3
        val mutable twoSided: uint32

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type BnCamera = // 88
        val mutable kind: uint32 // 0 pinhole, 1 thin-lens
        val mutable fovY: float32 // degrees, as PinholeCamera.FovY
        val mutable aspectRatio: float32
        val mutable aperture: float32
        val mutable focusDistance: float32
        val mutable pushForward: float32
        val mutable cameraToWorld: Matrix4x4

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type BnSceneDesc = // 264: eleven (pointer, count) pairs of 16 bytes, the camera at offset 172
        val mutable tlasNodes: nativeint
        val mutable tlasNodeCount: uint32
        val mutable instances: nativeint
        val mutable instanceCount: uint32
        val mutable lightInstances: nativeint
        val mutable lightInstanceCount: uint32
        val mutable meshes: nativeint
        val mutable meshCount: uint32
        val mutable vertices: nativeint
        val mutable vertexCount: uint32
        val mutable triangles: nativeint
        val mutable triangleCount: uint32
        val mutable blasNodes: nativeint
        val mutable blasNodeCount: uint32
        val mutable alias: nativeint
        val mutable aliasCount: uint32
        val mutable sphereRadii: nativeint
        val mutable sphereCount: uint32
        val mutable materials: nativeint
        val mutable materialCount: uint32
        val mutable lights: nativeint
        val mutable lightCount: uint32
        val mutable camera: BnCamera

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type BnRenderParams = // 64
        val mutable width: int
        val mutable height: int
        val mutable spp: int
        val mutable maxDepth: int
        val mutable rrDepth: int
        val mutable frameId: int
        val mutable sampleBegin: int
        val mutable sampleEnd: int
        val mutable x0: int
        val mutable y0: int
        val mutable x1: int
        val mutable y1: int
        val mutable flags: uint32
        val mutable interleaveCount: int
        val mutable interleaveIndex: int
        val mutable integrator: int // 0 path tracing, 1 direct, 2 normal

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type BnStats = // 80
        val mutable paths: uint64
        val mutable extendRays: uint64
        val mutable shadowRays: uint64
        val mutable shadowRaysRef: uint64
        val mutable kernelLaunches: uint64
        val mutable gpuMs: float
        val mutable extendMs: float
        val mutable shadeMs: float
        val mutable shadowMs: float
        val mutable otherMs: float

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type BnMltParams = // 56
        val mutable width: int
        val mutable height: int
        val mutable mutationsPerPixel: int
        val mutable maxDepth: int
        val mutable rrDepth: int
        val mutable frameId: int
        val mutable nBootstrap: int
        val mutable nChains: int
        val mutable strategy: int // 0 Gaussian (p0 = sigma), 1 Kelemen (p0 = epsMin, p1 = epsMax)
        val mutable p0: float32
        val mutable p1: float32
        val mutable largeStepProb: float32
        val mutable chainBegin: int
        val mutable chainEnd: int

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type BnMltStats = // 48
        val mutable b: float32
        val mutable reserved: uint32
        val mutable accepted: uint64
        val mutable proposed: uint64
        val mutable rays: uint64
        val mutable bootstrapMs: float
        val mutable chainsMs: float

    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int bn_device_count()

    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern nativeint bn_last_error()

    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int bn_scene_create(BnSceneDesc& desc, int device, nativeint& scene)

    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern void bn_scene_destroy(nativeint scene)

    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int bn_render(nativeint scene, BnRenderParams& p, nativeint filmRgb, BnStats& stats)

    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int bn_render_pssmlt(nativeint scene, BnMltParams& p, nativeint filmRgb, BnMltStats& stats)

    /// Several devices behind one call (include/barnacle_b200.h): the scene is flattened once and uploaded to every listed
    /// device; bn_render_multi shares the (pixel, sampleId) space out inside the call — Integrator.Render's own parallelism
    /// (Integrator.fs:46-55) on the GPUs of the box instead of the TPL pool — and combines the films over NVLink on devices[0].
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int bn_multi_scene_create(BnSceneDesc& desc, nativeint devices, int nDevices, nativeint& scene)

    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern void bn_multi_scene_destroy(nativeint scene)

    /// partition: 0 auto (sample split when there are at least as many samples as devices, else tile rows), 1 sample, 2 tile
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int bn_render_multi(nativeint scene, BnRenderParams& p, int partition, nativeint filmRgb, BnStats& stats)

    /// Parity entry points (integration/fsharp/ParityDump.fs): fixed ray batches and per-path radiance.
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int bn_trace(nativeint scene, nativeint rays, uint64 n, int anyHit, nativeint hits)

    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int bn_render_radiance(nativeint scene, BnRenderParams& p, nativeint radiance)

    /// BVHNode.Build on the device (optional, "next" row N2): same nodes and permutation as Util/BVH.fs:239-247, byte for byte.
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int bn_bvh_build(int device, nativeint boxes, uint32 n, nativeint nodes, uint32 maxNodes, nativeint perm, nativeint ms)

    /// The reference's error convention is `failwith` (Base/LightSampler.fs:8-9, Loader.fs:204 ...).
    let check (rc: int) =
        if rc <> 0 then
            failwith (Marshal.PtrToStringUTF8(bn_last_error ()))

/// Implemented by the GPU integrators so that Scene.Render can hand over what
/// IntegratorBase.Render's signature does not carry: the instance array.
type IGpuIntegrator =
    /// Scene.Traverse(t)'s result in its ORIGINAL order (a copy taken before BVHAggregate permutes the
    /// array in place, Extensions/Aggregate/BVH.fs:9 + Util/BVH.fs:244-246).
    abstract member Instances: PrimitiveInstance array with get, set

/// Drop-in for BVHNode.Build (Util/BVH.fs:239-247) that runs the binned-SAH build on the GPU: permutes `xs` in place,
/// returns the preorder node array.  Boxes must be finite (anything else fails loudly; keep the host builder for those).
[<AbstractClass; Sealed>]
type GpuBvh =
    /// A tupled static member like BVHNode.Build itself: Span is byref-like and cannot be a curried argument.
    static member Build(device: int, xs: 'a Span, f: 'a -> AxisAlignedBoundingBox) : BVHNode array =
            let xs' = xs.ToArray()
            let boxes = Array.zeroCreate<float32> (6 * xs'.Length)
            for i = 0 to xs'.Length - 1 do
                let b = f xs'[i]
                boxes[6 * i] <- b.pMin.X
                boxes[6 * i + 1] <- b.pMin.Y
                boxes[6 * i + 2] <- b.pMin.Z
                boxes[6 * i + 3] <- b.pMax.X
                boxes[6 * i + 4] <- b.pMax.Y
                boxes[6 * i + 5] <- b.pMax.Z
            let nodes = Array.zeroCreate<BVHNode> (max 1 (2 * xs'.Length))
            let perm = Array.zeroCreate<uint32> xs'.Length
            use pBoxes = fixed boxes
            use pNodes = fixed nodes // pinned, not marshalled: BVHNode holds a bool, which the marshaller would widen
            use pPerm = fixed perm
            let count =
                Native.bn_bvh_build (device, NativePtr.toNativeInt pBoxes, uint32 xs'.Length, NativePtr.toNativeInt pNodes,
                                     uint32 nodes.Length, NativePtr.toNativeInt pPerm, 0n)
            if count < 0 then
                failwith (Marshal.PtrToStringUTF8(Native.bn_last_error ()))
            for i = 0 to xs'.Length - 1 do
                xs[i] <- xs'[int perm[i]] // Util/BVH.fs:244-246
            Array.sub nodes 0 count

/// Flattens the managed object graph into the POD arrays of BnSceneDesc and keeps them pinned for the
/// duration of `body`.  Reads public members only.
module GpuScene =
    let private idOf (table: Dictionary<'k, int>) (key: 'k) (add: unit -> unit) =
        match table.TryGetValue key with
        | true, id -> id
        | _ ->
            let id = table.Count
            table[key] <- id
            add ()
            id

    /// `original`: Scene.Traverse(t) in its original order.  BVHAggregate.BVHNodes is private
    /// (Extensions/Aggregate/BVH.fs:9), so the TLAS is rebuilt here with the public, deterministic
    /// BVHNode.Build on a copy of the SAME input: same nodes, same permutation as the CPU aggregate.
    let withFlattenedScene (camera: CameraBase) (lightSampler: LightSamplerBase) (original: PrimitiveInstance array)
                           (body: Native.BnSceneDesc -> unit) =
        if original.Length = 0 then
            failwith "GpuScene: Instances not set (Scene.Render must assign IGpuIntegrator.Instances)"
        let ordered = Array.copy original
        let tlas = BVHNode.Build(ordered.AsSpan(), _.Bounds)

        let meshIds = Dictionary<MeshPrimitive, int>(HashIdentity.Reference)
        let sphereIds = Dictionary<SpherePrimitive, int>(HashIdentity.Reference)
        let materialIds = Dictionary<MaterialBase, int>(HashIdentity.Reference)
        let lightIds = Dictionary<LightBase, int>(HashIdentity.Reference)
        let meshes = ResizeArray<Native.BnMesh>()
        let vertices = ResizeArray<Vector3>()
        let triangles = ResizeArray<TriangleIndex>() // 3 x int32, BLAS order, local to the mesh's vertex slice
        let blasNodes = ResizeArray<BVHNode>() // 32-B explicit layout (Util/BVH.fs:52-73): passed as is
        let alias = ResizeArray<Entry>() // {alias; prob; pdf}, 12 B (Util/AliasTable.fs:7-12): passed as is
        let radii = ResizeArray<float32>()
        let materials = ResizeArray<Native.BnMaterial>()
        let lights = ResizeArray<Native.BnLight>()
        let lightInstances = ResizeArray<uint32>()

        let meshId (mesh: MeshPrimitive) =
            idOf meshIds mesh (fun () ->
                let mutable m = Native.BnMesh()
                m.vertexOffset <- uint32 vertices.Count
                m.vertexCount <- uint32 mesh.Vertices.Length
                m.triOffset <- uint32 triangles.Count
                m.triCount <- uint32 mesh.TriangleIndices.Length
                m.nodeOffset <- uint32 blasNodes.Count
                m.nodeCount <- uint32 mesh.BVHNodes.Length
                m.aliasOffset <- uint32 alias.Count
                vertices.AddRange mesh.Vertices
                triangles.AddRange mesh.TriangleIndices
                blasNodes.AddRange mesh.BVHNodes
                alias.AddRange mesh.AliasTable.Table
                meshes.Add m)

        let sphereId (sphere: SpherePrimitive) =
            idOf sphereIds sphere (fun () -> radii.Add sphere.Radius)

        let materialId (material: MaterialBase) =
            idOf materialIds material (fun () ->
                let mutable m = Native.BnMaterial()
                match material with
                | :? Lambertian as x ->
                    m.kind <- 0u
                    m.baseColor <- x.BaseColor
                | :? MirrorMaterial as x ->
                    m.kind <- 1u
                    m.baseColor <- x.BaseColor
                | :? DielectricMaterial as x ->
                    m.kind <- 2u
                    m.baseColor <- x.BaseColor
                    m.p0 <- x.IOR
                | :? PBRMaterial as x ->
                    m.kind <- 3u
                    m.baseColor <- x.BaseColor
                    m.p0 <- x.Metallic // already clamped (PBR.fs:12)
                    m.p1 <- x.Alpha // already max(roughness^2, 1e-3) (PBR.fs:11)
                | _ -> failwith $"GpuScene: unsupported material type: {material.GetType().Name}"
                materials.Add m)

        let lightId (light: LightBase) =
            idOf lightIds light (fun () ->
                let mutable l = Native.BnLight()
                match light with
                | :? DiffuseLight as x ->
                    l.emission <- x.Emission
                    l.twoSided <- if x.TwoSided then 1u else 0u
                | _ -> failwith $"GpuScene: unsupported light type: {light.GetType().Name}"
                lights.Add l)

        let instances =
            ordered
            |> Array.mapi (fun i inst ->
                let mutable r = Native.BnInstance()
                match inst with
                | :? MeshInstance as x ->
                    r.primKind <- 0u
                    r.primId <- uint32 (meshId x.Mesh)
                | :? SphereInstance as x ->
                    r.primKind <- 1u
                    r.primId <- uint32 (sphereId x.Sphere)
                | _ -> failwith $"GpuScene: unsupported primitive instance type: {inst.GetType().Name}"
                r.materialId <- if inst.HasMaterial then materialId inst.Material else -1
                r.lightId <- if inst.HasLight then lightId inst.Light else -1
                r.objectToWorld <- inst.ObjectToWorld
                r.worldToObject <- inst.WorldToObject
                r.boundsMin <- inst.Bounds.pMin
                r.boundsMax <- inst.Bounds.pMax
                if inst.HasLight then
                    lightInstances.Add(uint32 i)
                r)

        // LightSamplerBase.Instances (Base/LightSampler.fs:7,10) is the HasLight filter of the array the
        // CPU aggregate permuted: it must be the very same sequence, or the two TLAS builds disagree.
        let cpuLights = lightSampler.Instances
        if cpuLights.Length <> lightInstances.Count
           || not (Seq.forall2 (fun (a: PrimitiveInstance) (k: uint32) -> obj.ReferenceEquals(a, ordered[int k])) cpuLights lightInstances) then
            failwith "GpuScene: light order differs from the CPU light sampler's (Instances must be the unpermuted Scene.Traverse result)"

        let mutable cam = Native.BnCamera()
        match camera with
        | :? ThinLensCamera as c -> // before PinholeCamera: it inherits from it (ThinLens.fs:8-9)
            cam.kind <- 1u
            cam.fovY <- c.FovY
            cam.aspectRatio <- c.AspectRatio
            cam.aperture <- c.Aperture
            cam.focusDistance <- c.FocusDistance
        | :? PinholeCamera as c ->
            cam.kind <- 0u
            cam.fovY <- c.FovY
            cam.aspectRatio <- c.AspectRatio
            cam.focusDistance <- 1f
        | _ -> failwith $"GpuScene: unsupported camera type: {camera.GetType().Name}"
        cam.pushForward <- camera.PushForward
        cam.cameraToWorld <- camera.CameraToWorld

        let meshArr, vertArr, triArr, nodeArr, aliasArr = meshes.ToArray(), vertices.ToArray(), triangles.ToArray(), blasNodes.ToArray(), alias.ToArray()
        let radiiArr, matArr, lightArr, lightInstArr = radii.ToArray(), materials.ToArray(), lights.ToArray(), lightInstances.ToArray()
        // `fixed` pins for the rest of this function; an empty array pins to a null pointer (count 0).
        use pTlas = fixed tlas
        use pInst = fixed instances
        use pLightInst = fixed lightInstArr
        use pMesh = fixed meshArr
        use pVert = fixed vertArr
        use pTri = fixed triArr
        use pNode = fixed nodeArr
        use pAlias = fixed aliasArr
        use pRadii = fixed radiiArr
        use pMat = fixed matArr
        use pLight = fixed lightArr

        let mutable desc = Native.BnSceneDesc()
        desc.tlasNodes <- NativePtr.toNativeInt pTlas
        desc.tlasNodeCount <- uint32 tlas.Length
        desc.instances <- NativePtr.toNativeInt pInst
        desc.instanceCount <- uint32 instances.Length
        desc.lightInstances <- NativePtr.toNativeInt pLightInst
        desc.lightInstanceCount <- uint32 lightInstArr.Length
        desc.meshes <- NativePtr.toNativeInt pMesh
        desc.meshCount <- uint32 meshArr.Length
        desc.vertices <- NativePtr.toNativeInt pVert
        desc.vertexCount <- uint32 vertArr.Length
        desc.triangles <- NativePtr.toNativeInt pTri
        desc.triangleCount <- uint32 triArr.Length
        desc.blasNodes <- NativePtr.toNativeInt pNode
        desc.blasNodeCount <- uint32 nodeArr.Length
        desc.alias <- NativePtr.toNativeInt pAlias
        desc.aliasCount <- uint32 aliasArr.Length
        desc.sphereRadii <- NativePtr.toNativeInt pRadii
        desc.sphereCount <- uint32 radiiArr.Length
        desc.materials <- NativePtr.toNativeInt pMat
        desc.materialCount <- uint32 matArr.Length
        desc.lights <- NativePtr.toNativeInt pLight
        desc.lightCount <- uint32 lightArr.Length
        desc.camera <- cam

        body desc // every array above stays pinned until `body` returns; the library copies and keeps no host pointers

    /// One device: bn_scene_create / bn_scene_destroy around `body`.
    let withDeviceScene (camera: CameraBase) (lightSampler: LightSamplerBase) (original: PrimitiveInstance array) (device: int)
                        (body: nativeint -> unit) =
        withFlattenedScene camera lightSampler original (fun desc ->
            let mutable d = desc
            let mutable scene = 0n
            Native.check (Native.bn_scene_create (&d, device, &scene))
            try
                body scene
            finally
                Native.bn_scene_destroy scene)

    /// Several devices: the scene is flattened ONCE (above) and uploaded to every device by bn_multi_scene_create.
    let withMultiDeviceScene (camera: CameraBase) (lightSampler: LightSamplerBase) (original: PrimitiveInstance array) (devices: int array)
                             (body: nativeint -> unit) =
        let pin = GCHandle.Alloc(devices, GCHandleType.Pinned) // (`fixed` is not available inside the closure below)
        try
            withFlattenedScene camera lightSampler original (fun desc ->
                let mutable d = desc
                let mutable scene = 0n
                Native.check (Native.bn_multi_scene_create (&d, pin.AddrOfPinnedObject(), devices.Length, &scene))
                try
                    body scene
                finally
                    Native.bn_multi_scene_destroy scene)
        finally
            pin.Free()

/// kind: 0 = PathTracingIntegrator.Li (PathTracing.fs:14-81), 1 = DirectIntegrator.Li (Direct.fs:10-40),
/// 2 = NormalIntegrator.Li (Normal.fs:10-17) — all under ProgressiveIntegrator.Render (Integrator.fs:22-55).
[<Sealed>]
type GpuProgressiveIntegrator(spp: int, maxDepth: int, rrDepth: int, kind: int) =
    inherit ProgressiveIntegrator(spp)
    let mutable instances: PrimitiveInstance array = [||]
    member val Device = 0 with get, set
    /// More than one entry: the frame is rendered by all of them inside ONE bn_render_multi call (e.g. [| 0 .. 7 |] on the
    /// 8 x B200 box) — the GPU counterpart of Parallel.ForEach over the tiles (Integrator.fs:46-54).
    member val Devices: int array = [||] with get, set
    /// 0 auto, 1 sample split, 2 tile split (BN_PARTITION_*)
    member val Partition = 0 with get, set
    member val LastStats = Native.BnStats() with get, set
    member this.MaxDepth = maxDepth
    member this.RRDepth = rrDepth

    interface IGpuIntegrator with
        member _.Instances
            with get () = instances
            and set v = instances <- v

    override this.Render(camera, film, _aggregate, lightSampler) =
        use pixels = fixed film.Pixels // Vector3[W*H] == float[3*W*H], already Y-flipped (Film.fs:41-46); pinned for the whole call
        let filmRgb = NativePtr.toNativeInt pixels
        let mutable p = Native.BnRenderParams()
        p.width <- film.ImageWidth
        p.height <- film.ImageHeight
        p.spp <- this.SamplePerPixel
        p.maxDepth <- maxDepth
        p.rrDepth <- rrDepth
        p.frameId <- this.FrameId
        p.sampleBegin <- 0
        p.sampleEnd <- this.SamplePerPixel
        p.x0 <- 0
        p.y0 <- 0
        p.x1 <- film.ImageWidth
        p.y1 <- film.ImageHeight
        p.flags <- 0u
        p.interleaveCount <- 1
        p.interleaveIndex <- 0
        p.integrator <- kind
        let p0 = p
        if this.Devices.Length > 1 then
            GpuScene.withMultiDeviceScene camera lightSampler instances this.Devices (fun scene ->
                let mutable q = p0
                let mutable stats = Native.BnStats()
                Native.check (Native.bn_render_multi (scene, &q, this.Partition, filmRgb, &stats))
                this.LastStats <- stats)
        else
            let device = if this.Devices.Length = 1 then this.Devices[0] else this.Device
            GpuScene.withDeviceScene camera lightSampler instances device (fun scene ->
                let mutable q = p0
                let mutable stats = Native.BnStats()
                Native.check (Native.bn_render (scene, &q, filmRgb, &stats))
                this.LastStats <- stats)
        this.FrameId <- this.FrameId + 1 // Integrator.fs:55

/// PSSMLTIntegrator.Render (PSSMLT.fs:379-414) on the device.
[<Sealed>]
type GpuPSSMLTIntegrator
    (maxDepth: int, rrDepth: int, nBootstrap: int, nChains: int, mutationPerPixel: int, strategy: MutationStrategy, largeStepProb: float32) =
    inherit ProgressiveIntegrator(mutationPerPixel) // as PSSMLTIntegrator does (PSSMLT.fs:155): FrameId seeds the samplers
    let mutable instances: PrimitiveInstance array = [||]
    member val Device = 0 with get, set
    member val B = 0f with get, set
    member val AcceptedMutationCount = 0L with get, set
    member val ProposedMutationCount = 0L with get, set

    interface IGpuIntegrator with
        member _.Instances
            with get () = instances
            and set v = instances <- v

    override this.Render(camera, film, _aggregate, lightSampler) = // FrameId is not advanced (PSSMLT.fs:379-414 never does)
        use pixels = fixed film.Pixels // accumulated INTO, like Film.Accumulate; Scene.Render cleared it
        let filmRgb = NativePtr.toNativeInt pixels
        GpuScene.withDeviceScene camera lightSampler instances this.Device (fun scene ->
            let mutable p = Native.BnMltParams()
            p.width <- film.ImageWidth
            p.height <- film.ImageHeight
            p.mutationsPerPixel <- mutationPerPixel
            p.maxDepth <- maxDepth
            p.rrDepth <- rrDepth
            p.frameId <- this.FrameId // Sampler(FrameId, bootstrapId) / Sampler(FrameId, chainId), PSSMLT.fs:252,286
            p.nBootstrap <- nBootstrap
            p.nChains <- nChains
            match strategy with
            | Gaussian sigma ->
                p.strategy <- 0
                p.p0 <- sigma
            | Kelemen(epsMin, epsMax) ->
                p.strategy <- 1
                p.p0 <- epsMin
                p.p1 <- epsMax
            p.largeStepProb <- largeStepProb
            p.chainBegin <- 0
            p.chainEnd <- nChains
            let mutable stats = Native.BnMltStats()
            Native.check (Native.bn_render_pssmlt (scene, &p, filmRgb, &stats))
            this.B <- stats.b
            this.AcceptedMutationCount <- int64 stats.accepted
            this.ProposedMutationCount <- int64 stats.proposed
            if this.B = 0f then
                printfn "Warning: all bootstrap samples are zero, exiting..." // PSSMLT.fs:396-397
            else
                printfn $"Accepted mutation count: %d{this.AcceptedMutationCount}"
                printfn $"Proposed mutation count: %d{this.ProposedMutationCount}"
                printfn $"Acceptance rate: %f{float this.AcceptedMutationCount / float this.ProposedMutationCount}")

// ------------------------------------------------------------------------------------------------
// The two edits in existing reference files.
//
// (1) Extensions/Scene/Loader.fs, IntegratorInfo.ToIntegrator (:185-204) — four more arms, placed before
//     the catch-all; the pssmlt arm reuses the bindings the "pssmlt" arm computes:
//
//         (optionally `GpuProgressiveIntegrator(...) |> fun g -> g.Devices <- [| 0 .. Native.bn_device_count () - 1 |]; g` to use every GPU of the box)
//         | "gpu-path-tracing" -> GpuProgressiveIntegrator(spp, maxDepth, rrDepth, 0)
//         | "gpu-direct" -> GpuProgressiveIntegrator(spp, maxDepth, rrDepth, 1)
//         | "gpu-normal" -> GpuProgressiveIntegrator(spp, maxDepth, rrDepth, 2)
//         | "gpu-pssmlt" ->
//             ... (nBootstrap, nChains, mutationStrategy, largeStepProb exactly as in the "pssmlt" arm)
//             GpuPSSMLTIntegrator(maxDepth, rrDepth, nBootstrap, nChains, spp, mutationStrategy, largeStepProb)
//
// (2) Extensions/Scene/Render.fs:12-13 — hand the unpermuted instance array to a GPU integrator before
//     BVHAggregate permutes it in place:
//
//         let instances = this.Traverse(t)
//         match box this.Integrator with
//         | :? IGpuIntegrator as gpu -> gpu.Instances <- Array.copy instances
//         | _ -> ()
//         let aggregate = BVHAggregate(instances)
//         let lightSampler = UniformLightSampler(instances)
//
// With both, `dotnet run -i scene.json -o img` on a scene whose integrator type is "gpu-path-tracing"
// renders through the CUDA library inside the reference's own Stopwatch region (Render.fs:15-17), and
// Film.Save proceeds unchanged.
// ------------------------------------------------------------------------------------------------
