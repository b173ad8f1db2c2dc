// ParityDump.fs — the pin that cannot be made in the build image: the REFERENCE's own arithmetic on the repo's fixed
// inputs.  STATUS: source only (no `dotnet` in the image), never compiled.  On a box with .NET 9:
//
//   * copy this file to Extensions/Integrator/ParityDump.fs and add it to Barnacle.fsproj AFTER Extensions\Scene\Render.fs
//   * Program.fs: before the normal argument handling,
//         if argv.Length >= 1 && argv[0] = "--parity-dump" then exit (ParityDump.main argv[1..])
//   * from the reference's root (so that `Asset/...` URIs resolve; copy this repo's scenes/ there for the derived scenes):
//         dotnet run -c Release -- --parity-dump Asset/cbox.json           <repo>/tests/golden/dotnet/cbox_ref
//         dotnet run -c Release -- --parity-dump scenes/cbox_bunny.json    <repo>/tests/golden/dotnet/cbox_bunny
//     Each call reads  <prefix>.rays.bin  (committed: n x {origin f32x3, direction f32x3, tmax f32}, 28 B, little endian — the
//     layout of BnRay) and  <prefix>.window.txt  ("width height spp x0 y0 x1 y1 maxDepth rrDepth"), and writes
//         <prefix>.hits.bin      n x {t f32, u f32, v f32, instance i32, primitive i32}  (BnHit, 20 B)
//         <prefix>.anyhit.bin    n x u8: PrimitiveAggregate.Intersect/2 with the ray's tmax
//         <prefix>.radiance.bin  [sample][y - y0][x - x0][3] f32: Li * ReciprocalEstimate(pdf) of PathTracingIntegrator for the
//                                window (the layout of bn_render_radiance)
//         <prefix>.meta.txt      runtime version, Vector.IsHardwareAccelerated, Fma.IsSupported
//   * commit the four output files; tests/test_dotnet_dump.py then compares them with the C++ restatement (both BN_NET9_FMA
//     conventions) and, on a GPU box, with the CUDA path.  Hits are expected to be bit-identical for ONE of the two
//     conventions — that is the pin; radiance differs by libm-vs-fixed-polynomial transcendentals and MathF.ReciprocalEstimate
//     (SURVEY Q11) and is compared statistically.
//
// What is called is the reference's own public API only: Scene.Load, Scene.Traverse, BVHAggregate, UniformLightSampler,
// PathTracingIntegrator.Li, CameraBase.GeneratePrimaryRay, Sampler — no code of this repo runs.
namespace Barnacle.Extensions.Integrator

open System
open System.IO
open System.Numerics
open Barnacle.Base
open Barnacle.Extensions.Aggregate
open Barnacle.Extensions.LightSampler
open Barnacle.Extensions.Primitive
open Barnacle.Extensions.Scene

module ParityDump =
    let private readRays (path: string) =
        let bytes = File.ReadAllBytes path
        let n = bytes.Length / 28
        Array.init n (fun i ->
            let f k = BitConverter.ToSingle(bytes, 28 * i + 4 * k)
            struct (Ray(Vector3(f 0, f 1, f 2), Vector3(f 3, f 4, f 5)), f 6))

    /// Closest hit of every ray through the reference's TLAS + BLAS walk (Extensions/Aggregate/BVH.fs:37-58).
    /// `instance` = index in the TLAS-ordered instance array (the order BVHAggregate's constructor leaves `instances` in,
    /// Util/BVH.fs:244-246); `primitive` = LocalGeometry.tag as MeshPrimitive.Intersect sets it (Mesh.fs:232), i.e. BEFORE
    /// LocalGeometry.Transform resets it to 0 (Primitive.fs:57-58): the hit instance's primitive is asked again in object space —
    /// the same walk over the same tree finds the same closest triangle, ties included (first visited wins, and the visiting
    /// order does not depend on the initial t).
    let dumpHits (instances: PrimitiveInstance array) (aggregate: PrimitiveAggregate) (rays: struct (Ray * float32) array) (prefix: string) =
        use hits = new BinaryWriter(File.Create(prefix + ".hits.bin"))
        use anyhit = new BinaryWriter(File.Create(prefix + ".anyhit.bin"))
        for struct (ray, tmax) in rays do
            let mutable interaction = Unchecked.defaultof<Interaction>
            let mutable t = tmax
            let ray' = ray
            if aggregate.Intersect(&ray', &interaction, &t) then
                let inst = interaction.inst
                let index = Array.FindIndex(instances, fun x -> obj.ReferenceEquals(x, inst))
                let mutable tag = 0
                match inst with
                | :? MeshInstance ->
                    let objRay = Ray.Transform(&ray', inst.WorldToObject)
                    let mutable geom = Unchecked.defaultof<LocalGeometry>
                    let mutable t2 = tmax
                    if inst.Primitive.Intersect(&objRay, &geom, &t2) then
                        tag <- geom.tag
                | _ -> ()
                hits.Write t
                hits.Write interaction.UV.X
                hits.Write interaction.UV.Y
                hits.Write index
                hits.Write tag
            else
                hits.Write t // unchanged tmax
                hits.Write 0f
                hits.Write 0f
                hits.Write -1
                hits.Write -1
            anyhit.Write(if aggregate.Intersect(&ray', tmax) then 1uy else 0uy)

    /// RenderTile's sample loop (Base/Integrator.fs:34-42) without the accumulation: one radiance per (sampleId, pixel).
    let dumpRadiance (scene: Scene) (aggregate: PrimitiveAggregate) (lightSampler: LightSamplerBase) (w: int) (h: int) (spp: int)
                     (x0: int, y0: int, x1: int, y1: int) (maxDepth: int) (rrDepth: int) (prefix: string) =
        let integrator = PathTracingIntegrator(spp, maxDepth, rrDepth)
        use out = new BinaryWriter(File.Create(prefix + ".radiance.bin"))
        for sampleId = 0 to spp - 1 do
            for y = y0 to y1 - 1 do
                for x = x0 to x1 - 1 do
                    let mutable sampler = Sampler(uint x, uint y, uint (integrator.FrameId * spp + sampleId))
                    let struct (ray, pdf) =
                        scene.Camera.GeneratePrimaryRay(struct (w, h), struct (x, y), sampler.Next2D(), sampler.Next2D())
                    let radiance = integrator.Li(&ray, aggregate, lightSampler, &sampler) * MathF.ReciprocalEstimate(pdf)
                    out.Write radiance.X
                    out.Write radiance.Y
                    out.Write radiance.Z

    let main (argv: string array) =
        if argv.Length <> 2 then
            eprintfn "usage: --parity-dump <scene.json> <prefix>   (reads <prefix>.rays.bin and <prefix>.window.txt)"
            2
        else
            let scene = Scene.Load argv[0]
            let prefix = argv[1]
            let instances = scene.Traverse(0f) // Render.fs:12
            let aggregate = BVHAggregate(instances) // permutes `instances` into TLAS order (Render.fs:13)
            let lightSampler = UniformLightSampler(instances) // (Render.fs:14)
            dumpHits instances aggregate (readRays (prefix + ".rays.bin")) prefix
            let win = File.ReadAllText(prefix + ".window.txt").Split([| ' '; '\n'; '\r'; '\t' |], StringSplitOptions.RemoveEmptyEntries) |> Array.map int
            dumpRadiance scene aggregate lightSampler win[0] win[1] win[2] (win[3], win[4], win[5], win[6]) win[7] win[8] prefix
            File.WriteAllText(
                prefix + ".meta.txt",
                $"runtime {Environment.Version}\nVector.IsHardwareAccelerated {Vector.IsHardwareAccelerated}\n"
                + $"Fma.IsSupported {System.Runtime.Intrinsics.X86.Fma.IsSupported}\nscene {argv[0]}\n")
            printfn $"parity dump written to {prefix}.*"
            0
