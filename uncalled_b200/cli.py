"""`uncalled index` and `uncalled map` (reference scripts/uncalled:38-78,127-167 with the options of
uncalled/args.py:87-161,218-286) on this package: same sub-commands, option names, defaults, stderr progress
lines and PAF output.  `python -m uncalled_b200 map <prefix> <fast5s...>`.  `--device` picks the GPU of this
process; under `torchrun --nproc-per-node N -m uncalled_b200 map ...` every rank maps its share of the fast5 files on
its own GPU (reads are independent: no collective)."""
import argparse
import os
import sys
import time

MAX_SLEEP = 0.01


def get_parser(conf):
    from . import index_params as IP
    D = IP.DEFAULTS
    fmt = argparse.ArgumentDefaultsHelpFormatter
    parser = argparse.ArgumentParser(prog="uncalled_b200", description="Rapidly maps raw nanopore signal to DNA references",
                                     formatter_class=fmt)
    sp = parser.add_subparsers(dest="subcmd")
    p = sp.add_parser("index", help="Builds the UNCALLED index of a FASTA reference", formatter_class=fmt)
    p.add_argument("fasta_filename", type=str, help="FASTA file to index")
    p.add_argument("-o", "--bwa-prefix", type=str, default=None, help="Index output prefix. Will use input fasta filename by default")
    p.add_argument("-s", "--max-sample-dist", type=int, default=D["max_sample_dist"], help="Maximum average sampling distance between reference alignments.")
    p.add_argument("--min-samples", type=int, default=D["min_samples"], help="Minimum number of alignments to produce")
    p.add_argument("--max-samples", type=int, default=D["max_samples"], help="Maximum number of alignments to produce")
    p.add_argument("-k", "--kmer-len", type=int, default=D["kmer_len"], help="Model k-mer length")
    p.add_argument("-1", "--matchpr1", type=float, default=D["matchpr1"], help="Minimum event match probability")
    p.add_argument("-2", "--matchpr2", type=float, default=D["matchpr2"], help="Maximum event match probability")
    p.add_argument("-f", "--pathlen-percentile", type=float, default=D["pathlen_percentile"], help="")
    p.add_argument("-m", "--max-replen", type=int, default=D["max_replen"], help="")
    p.add_argument("--probs", type=str, default=None, help="Find parameters with specified target probabilites (comma separated)")
    p.add_argument("--speeds", type=str, default=None, help="Find parameters with specified speed coefficents (comma separated)")
    p.add_argument("--device", type=int, default=0, help="CUDA device")

    p = sp.add_parser("map", help="Map fast5 files to a DNA reference", formatter_class=fmt)
    p.add_argument("bwa_prefix", type=str, help="BWA prefix to mapping to. Must be processed by \"uncalled index\".")
    p.add_argument("-p", "--idx-preset", type=str, default=conf.idx_preset, help="Mapping mode")
    p.add_argument("fast5s", nargs="+", type=str, help="Reads to map. Can be a directory which will be recursively searched "
                   "for all files with the \".fast5\" extension, a text file containing one fast5 filename per line, or a "
                   "comma-separated list of fast5 file names.")
    p.add_argument("-r", "--recursive", action="store_true")
    p.add_argument("-l", "--read-list", type=str, default=None, help=type(conf).read_list.__doc__)
    p.add_argument("-n", "--max-reads", type=int, default=None, help=type(conf).max_reads.__doc__)
    p.add_argument("-t", "--threads", type=int, default=conf.threads, help="Number of host threads (fast5 decoding; the mapping runs on the GPU)")
    p.add_argument("--num-channels", type=int, default=conf.num_channels, help="Number of channels used in sequencing.")
    p.add_argument("-e", "--max-events", type=int, default=conf.max_events, help="Will give up on a read after this many events have been processed")
    p.add_argument("-c", "--max-chunks", type=int, default=conf.max_chunks, help="Will give up on a read after this many chunks have been processed.")
    p.add_argument("--chunk-time", type=float, default=1, required=False, help="Length of chunks in seconds")
    p.add_argument("--device", type=int, default=conf.device, help="CUDA device")
    p.add_argument("--batch-reads", type=int, default=conf.batch_reads, help="Reads per GPU batch")
    p.add_argument("--ordered", action="store_const", const=1, default=conf.ordered, help=type(conf).ordered.__doc__)
    p.add_argument("--exact-ties", action="store_const", const=1, default=conf.exact_ties, help=type(conf).exact_ties.__doc__)

    p = sp.add_parser("pafstats", help="Computes speed and accuracy of UNCALLED mappings.", formatter_class=fmt)
    p.add_argument("infile", type=str, help="PAF file output by UNCALLED")          # uncalled/pafstats.py:165-169
    p.add_argument("-n", "--max-reads", required=False, type=int, default=None, help="Will only look at first n reads if specified")
    p.add_argument("-r", "--ref-paf", required=False, type=str, default=None, help="Reference PAF file. Will output percent true/false "
                   "positives/negatives with respect to reference. Reads not mapped in reference PAF will be classified as NA.")
    return parser


def fast5_path(fname):
    if fname.startswith("#") or not fname.endswith("fast5"):
        return None
    path = os.path.abspath(fname)
    if not os.path.isfile(path):
        sys.stderr.write("Warning: \"%s\" is not a fast5 file.\n" % fname)
        return None
    return path


def load_fast5s(fast5s, recursive):
    """scripts/uncalled:88-118: directories (optionally recursive), .fast5 files, or text files of file names."""
    for path in fast5s:
        path = path.strip()
        if not os.path.exists(path):
            sys.stderr.write("Error: \"%s\" does not exist\n" % path)
            sys.exit(1)
        if os.path.isdir(path) and recursive:
            for root, _, files in os.walk(path):
                for fname in files:
                    yield fast5_path(os.path.join(root, fname))
        elif os.path.isdir(path):
            for fname in os.listdir(path):
                yield fast5_path(os.path.join(path, fname))
        elif path.endswith(".fast5"):
            yield fast5_path(path)
        else:
            with open(path) as infile:
                for line in infile:
                    yield fast5_path(line.strip())


def assert_exists(fname):
    if not os.path.exists(fname):
        sys.stderr.write("Error: '%s' does not exist\n" % fname)
        sys.exit(1)


def index_cmd(args):
    from . import _native as N
    from . import index as UI
    N.check(N.lib().unc_init(args.device))
    opts = {k: getattr(args, k) for k in ("max_sample_dist", "min_samples", "max_samples", "kmer_len", "matchpr1", "matchpr2",
                                          "pathlen_percentile", "max_replen")}
    UI.index_cmd(args.fasta_filename, args.bwa_prefix, probs=args.probs, speeds=args.speeds, **opts)


def map_cmd(conf, args, out=None):
    from .api import MapPool
    assert_exists(conf.bwa_prefix + ".bwt")
    assert_exists(conf.bwa_prefix + ".uncl")
    if len(conf.read_list) > 0:
        assert_exists(conf.read_list)
    # one process per GPU (torchrun / mpirun): rank r maps the files i with i mod WORLD_SIZE == r on GPU LOCAL_RANK
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if world > 1:
        conf.device = int(os.environ.get("LOCAL_RANK", rank))
    mapper = MapPool(conf)
    sys.stderr.write("Loading fast5s\n")
    from .shard import local_files
    files = [f for f in load_fast5s(args.fast5s, args.recursive) if f is not None]
    for fast5 in local_files(files, world, rank):
        mapper.add_fast5(fast5)
    sys.stderr.write("Mapping\n")
    sys.stderr.flush()
    try:
        while mapper.running():
            t0 = time.time()
            for p in mapper.update():
                p.print_paf(out)
            dt = time.time() - t0
            if dt < MAX_SLEEP:
                time.sleep(MAX_SLEEP - dt)
    except KeyboardInterrupt:
        pass
    sys.stderr.write("Finishing\n")
    mapper.stop()


def load_conf(argv):
    """uncalled/args.py:288-302: every parsed option whose name is a Conf attribute is set on the Conf."""
    from .api import Conf
    conf = Conf()
    parser = get_parser(conf)
    args = parser.parse_args(argv)
    for a, v in vars(args).items():
        if v is not None and not a.startswith("_") and hasattr(conf, a):
            setattr(conf, a, v)
    return parser, conf, args


def main(argv=None):
    parser, conf, args = load_conf(sys.argv[1:] if argv is None else argv)
    if args.subcmd == "index":
        index_cmd(args)
    elif args.subcmd == "map":
        map_cmd(conf, args)
    elif args.subcmd == "pafstats":
        from . import pafstats
        pafstats.run(args.infile, args.ref_paf, args.max_reads)
    else:
        parser.print_help()
    return 0
