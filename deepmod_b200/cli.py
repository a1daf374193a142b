"""``DeepMod.py detect`` command line, flag-for-flag (``bin/DeepMod.py:304-338``).

Only ``detect`` is provided: ``train`` and ``getfeatures`` are the reference's training side
and not part of the GPU hot path.
"""
import argparse
import os
import sys
from collections import defaultdict

OUTPUT_DEBUG, OUTPUT_INFO, OUTPUT_WARNING, OUTPUT_ERROR = 0, 1, 2, 3


def format_last_letter_of_folder(folder):
    if folder in (None, ""):
        return folder
    return folder if folder[-1] in "/\\" else folder + "/"


def build_parser():
    parser = argparse.ArgumentParser(
        description="Detect nucleotide modification from nanopore signals data (B200-native hot path).",
        formatter_class=argparse.RawTextHelpFormatter)
    sub = parser.add_subparsers()
    common = argparse.ArgumentParser(add_help=False)
    g = common.add_argument_group("Common options.")
    g.add_argument("--outLevel", type=int, choices=[OUTPUT_DEBUG, OUTPUT_INFO, OUTPUT_WARNING, OUTPUT_ERROR],
                   default=OUTPUT_WARNING, help="The level for output: 0 DEBUG, 1 INFO, 2 WARNING, 3 ERROR. Default: 2")
    g.add_argument("--wrkBase", help="The base folder of packed read batches (*.dmreads.npz).")
    g.add_argument("--FileID", default="mod", help="The unique string for output files. Default: 'mod'")
    g.add_argument("--outFolder", default="./mod_output", help="The folder for the results. Default: ./mod_output")
    g.add_argument("--recursive", type=int, default=1, choices=[0, 1], help="Recurse to find input files. Default: 1")
    g.add_argument("--threads", type=int, default=4, help="Accepted for compatibility; GPUs are selected by the launcher (one process per GPU).")
    g.add_argument("--files_per_thread", type=int, default=1000, help="Accepted for compatibility.")
    g.add_argument("--windowsize", type=int, default=21, help="The window size to extract features. Default: 21")
    g.add_argument("--alignStr", type=str, default="minimap2", choices=["bwa", "minimap2"], help="Accepted for compatibility (alignment is upstream of this path).")
    g.add_argument("--SignalGroup", type=str, default="simple", choices=["simple", "rundif"], help="Accepted for compatibility.")
    g.add_argument("--move", default=False, action="store_true", help="Accepted for compatibility.")
    det = sub.add_parser("detect", parents=[common], help="Detect modifications at a genomic scale",
                         formatter_class=argparse.RawTextHelpFormatter)
    det.add_argument("--Ref", help="The reference sequence (optional here: packed batches carry the contig table)")
    det.add_argument("--predDet", type=int, default=1, choices=[0, 1], help="pred first and then detect (1) or only detect (0). Default: 1")
    det.add_argument("--predpath", default=None, help="The file path of predictions for each fast5 file.")
    det.add_argument("--modfile", type=str, default=None, help="The path to load training model.")
    det.add_argument("--fnum", type=int, default=7, help="The number of features. Default: 7")
    det.add_argument("--hidden", type=int, default=100, help="The number of hidden node. Default: 100")
    det.add_argument("--basecall_1d", default="Basecall_1D_000", help="Accepted for compatibility.")
    det.add_argument("--basecall_2strand", default="BaseCalled_template", help="Accepted for compatibility.")
    det.add_argument("--region", default=None, help="The region of interest: for example, chr:1:100000;chr2:10000")
    det.add_argument("--ConUnk", default=True, choices=[False, True], help="Whether contain unknown chromosome")
    det.add_argument("--outputlayer", default="", choices=["", "sigmoid"], help="how to put activation function for output layer")
    det.add_argument("--Base", type=str, default="C", choices=["A", "C", "G", "T"], help="Interest of bases")
    det.add_argument("--mod_cluster", default=0, choices=[0, 1], help="1: CpG cluster effect; 0: not")
    det.add_argument("--precision", default="fp32", choices=["fp32", "f16", "bf16"],
                     help="fp32: parity path (<=1e-4 vs the reference graph); f16 / bf16: tcgen05 tensor-core path with fp16 / bf16 operands. Default: fp32")
    det.add_argument("--saveDetail", type=int, default=0, choices=[0, 1],
                     help="1: also write the per-read predictions and their index files (myDetect.py:716-782), the input of a later --predDet 0 run. Default: 0 (the per-position summary is accumulated on the GPU)")
    det.set_defaults(func=mDetect)
    return parser


def options_from_args(margs):
    """``mCommonParam`` + ``mDetect`` option assembly (bin/DeepMod.py:48-93, :99-160)."""
    err = ""
    mo = defaultdict()
    mo["outLevel"] = margs.outLevel
    mo["wrkBase"] = margs.wrkBase
    if mo["wrkBase"] is None:
        err += "\n\tThe input folder is None."
    mo["FileID"] = margs.FileID
    mo["outFolder"] = format_last_letter_of_folder(margs.outFolder)
    if mo["outFolder"] is not None and not os.path.isdir(mo["outFolder"]):
        try:
            os.makedirs(mo["outFolder"], exist_ok=True)
        except OSError:
            err += "\n\tThe output folder (%s) does not exist and cannot be created." % mo["outFolder"]
    mo["recursive"] = margs.recursive
    mo["files_per_thread"] = max(2, margs.files_per_thread)
    mo["threads"] = max(1, margs.threads)
    mo["windowsize"] = margs.windowsize
    if mo["windowsize"] < 1:
        err += "\n\tError windowsize could not be negative(%d)" % mo["windowsize"]
    mo["alignStr"] = margs.alignStr
    mo["SignalGroup"] = margs.SignalGroup
    mo["move"] = margs.move
    mo["basecall_1d"] = margs.basecall_1d
    mo["basecall_2strand"] = margs.basecall_2strand
    mo["ConUnk"] = margs.ConUnk
    mo["outputlayer"] = margs.outputlayer
    mo["Base"] = margs.Base
    mo["mod_cluster"] = margs.mod_cluster
    mo["precision"] = margs.precision
    mo["saveDetail"] = margs.saveDetail
    if mo["Base"] in ("", None):
        err += "\n\t Please provide a base of interest."
    mo["predDet"] = margs.predDet
    if mo["predDet"]:
        mo["Ref"] = margs.Ref
        if mo["Ref"] is not None and not os.path.isfile(mo["Ref"]):
            err += "\n\t reference file does not exist (%s)" % mo["Ref"]
        mo["fnum"] = margs.fnum
        mo["hidden"] = margs.hidden
        for k in ("fnum", "hidden"):
            if mo[k] < 1:
                err += "\n\tError %s could not be negative(%d)" % (k, mo[k])
        mo["modfile"] = margs.modfile
        if mo["modfile"] is None:
            err += "\n\tNo mod file is provided."
        elif not (os.path.isfile(mo["modfile"] + ".meta") or mo["modfile"].endswith(".npz") or os.path.isdir(mo["modfile"])):
            err += "\n\tThe meta file (%s) does not exist" % (mo["modfile"] + ".meta")
    else:
        mo["predpath"] = margs.predpath
        if mo["predpath"] is None or not os.path.isdir(mo["predpath"]):
            err += "\n\tThe predpath does not exist"
    mo["region"] = []
    if margs.region in (None, ""):
        mo["region"].append([None, None, None])
    else:
        for mr in margs.region.split(";"):
            sp = mr.split(":")
            mo["region"].append([sp[0], int(sp[1]) if len(sp) > 1 else None, int(sp[2]) if len(sp) > 2 else None])
    return mo, err


def printParameters(moptions):
    print("%30s: %s" % ("Current directory", os.getcwd()))
    for k in moptions.keys():
        print("%30s: %s" % (k, str(moptions[k])))
    sys.stdout.flush()


def mDetect(margs):
    moptions, err = options_from_args(margs)
    if int(os.environ.get("RANK", "0")) == 0:
        printParameters(moptions)
    if err:
        print("Please provide correct parameters" + err)
        sys.exit(1)
    from . import detect
    return detect.mDetect_manager(moptions)


def main(argv=None):
    parser = build_parser()
    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        parser.print_help()
        return None
    args = parser.parse_args(argv)
    if not hasattr(args, "func"):
        parser.print_help()
        return None
    return args.func(args)
