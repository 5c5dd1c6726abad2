"""Command line front end, the counterpart of the reference's `asrtool`
(cpp/bin/main.cpp:114-176):

    python -m asr_b200.asrtool --in point_cloud.ply --out mesh.ply

reads a PLY point cloud with normals (and optionally radii), runs
adaptivesurfacereconstruction.reconstruct_surface with the default parameters on the GPU and
writes the triangle mesh as PLY.  The network weights come from `model.pt` in the resource
directory (ASR_RESOURCE_DIR) like in the reference; `--model FILE` names another TorchScript
archive / state-dict file and `--random-weights SEED` runs with seeded random weights (no
meaningful surface; for smoke tests, since the released weights cannot be fetched offline)."""
import argparse
import sys


def main(argv=None):
    ap = argparse.ArgumentParser(prog="asrtool", description="usage: asrtool --in point_cloud.ply --out mesh.ply")
    ap.add_argument("--in", dest="inp", help="Input point cloud with normal information in PLY format.")
    ap.add_argument("--out", help="Output mesh in PLY format.")
    ap.add_argument("--version", action="store_true", help="Prints the version information")
    ap.add_argument("--third-party-notices", action="store_true", help="Prints third-party software notices")
    ap.add_argument("--model", help="weights file (TorchScript archive or state dict) instead of the resource dir")
    ap.add_argument("--random-weights", type=int, default=None, metavar="SEED")
    args = ap.parse_args(argv)
    import adaptivesurfacereconstruction as asr
    if args.version:
        print("asrtool version " + asr.get_version_str())
        return 0
    if args.third_party_notices:
        print(asr.get_third_party_notices())
        return 0
    if not args.inp or not args.out:
        ap.print_help()
        return 1
    from asr_b200 import model as _model
    from asr_b200 import plyio
    print("reading points")
    points, normals, radii = plyio.read_points(args.inp)
    print("%d / %d" % (len(points), len(points)))
    net = None
    if args.random_weights is not None:
        net = _model.seeded_weights(_model.UNet(5), seed=args.random_weights).cuda()
    elif args.model:
        net = _model.from_state_dict(_model.load_weights_file(args.model), 5)
    mesh = asr.reconstruct_surface(points, normals, radii if len(radii) else None, model=net)
    plyio.write_mesh(args.out, mesh["vertices"], mesh["triangles"])
    print("wrote %d vertices, %d triangles to %s" % (len(mesh["vertices"]), len(mesh["triangles"]), args.out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
