"""Command-line interface: same flags, me-PCR `K=V` translation, logging and exit codes as the reference's
`merpcr/cli.py:19-266` (argparse exits 2 on bad values; any failure in the run exits 1)."""
from __future__ import annotations

import argparse
import logging
import sys
from typing import List

from .engine import (DEFAULT_IUPAC_MODE, DEFAULT_MARGIN, DEFAULT_MISMATCHES, DEFAULT_PCR_SIZE, DEFAULT_THREADS,
                     DEFAULT_THREE_PRIME_MATCH, DEFAULT_WORDSIZE, MerPCR)

DEFAULT_MAX_STS_LINE_LENGTH = 1022
_MEPCR_FLAGS = {"M": "-M", "N": "-N", "W": "-W", "X": "-X", "T": "-T", "Q": "-Q", "Z": "-Z", "I": "-I", "S": "-S",
                "O": "-O"}


def convert_mepcr_arguments(args: List[str]) -> List[str]:
    """`M=50` -> `-M 50`; `P=...` (Mac priority) dropped; `-help` -> `--help` (cli.py:19-62)."""
    out: List[str] = []
    for arg in args:
        if len(arg) >= 3 and arg[1] == "=" and arg[0] in "MNWXTQZISOP":
            if arg[0] != "P":
                out.extend([_MEPCR_FLAGS[arg[0]], arg[2:]])
        elif arg == "-help":
            out.append("--help")
        else:
            out.append(arg)
    return out


def setup_logging(quiet: int, debug: bool) -> None:
    """cli.py:65-76."""
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s")
    logger = logging.getLogger("merpcr")
    if debug:
        logger.setLevel(logging.DEBUG)
    elif quiet == 0:
        logger.setLevel(logging.INFO)
    else:
        logger.setLevel(logging.WARNING)


def _bounded(name: str, lo, hi, fmt: str):
    def parse(value):
        ivalue = int(value)
        if (lo is not None and ivalue < lo) or (hi is not None and ivalue > hi):
            raise argparse.ArgumentTypeError(fmt.format(ivalue))
        return ivalue
    parse.__name__ = name
    return parse


margin_type = _bounded("margin_type", 0, 10000, "Margin must be between 0-10000, got {}")
mismatch_type = _bounded("mismatch_type", 0, 10, "Mismatches must be between 0-10, got {}")
wordsize_type = _bounded("wordsize_type", 3, 16, "Word size must be between 3-16, got {}")
threads_type = _bounded("threads_type", 1, None, "Threads must be > 0, got {}")
pcr_size_type = _bounded("pcr_size_type", 1, 10000, "PCR size must be between 1-10000, got {}")
sts_line_length_type = _bounded("sts_line_length_type", 1, None, "STS line length must be > 0, got {}")


# The reference's command line as data (cli.py:127-214): flag, long name, value parser, default, help text.  The help
# texts show the default the way the reference words them.
_OPTIONS = (
    ("-M", "--margin", margin_type, DEFAULT_MARGIN, "Margin (default: {d})"),
    ("-N", "--mismatches", mismatch_type, DEFAULT_MISMATCHES, "Number of mismatches allowed (default: {d})"),
    ("-W", "--wordsize", wordsize_type, DEFAULT_WORDSIZE, "Word size (default: {d})"),
    ("-T", "--threads", threads_type, DEFAULT_THREADS, "Number of threads (default: {d})"),
    ("-X", "--three-prime-match", int, DEFAULT_THREE_PRIME_MATCH,
     "Number of 3'-ward bases in which to disallow mismatches (default: {d})"),
    ("-O", "--output", str, None, "Output file name (default: stdout)"),
    ("-Q", "--quiet", (0, 1), 1, "Quiet flag (0=verbose, 1=quiet)"),
    ("-Z", "--default-pcr-size", pcr_size_type, DEFAULT_PCR_SIZE, "Default PCR size (default: {d})"),
    ("-I", "--iupac", (0, 1), DEFAULT_IUPAC_MODE,
     "IUPAC flag (0=don't honor IUPAC ambiguity symbols, 1=honor IUPAC symbols)"),
    ("-S", "--max-sts-line-length", sts_line_length_type, DEFAULT_MAX_STS_LINE_LENGTH,
     "Max. line length for the STS file (default: {d})"),
)


def create_parser() -> argparse.ArgumentParser:
    """The reference's parser (cli.py:127-214) plus the two switches of this implementation."""
    ap = argparse.ArgumentParser(description="merPCR - Modern Electronic Rapid PCR",
                                 formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    for name, text in (("sts_file", "STS file (tab-delimited)"), ("fasta_file", "FASTA sequence file")):
        ap.add_argument(name, type=str, help=text)
    for short, long_name, kind, default, text in _OPTIONS:
        kw = dict(default=default, help=text.format(d=default))
        if isinstance(kind, tuple):          # a closed set of integer values
            kw.update(type=int, choices=list(kind))
        else:
            kw["type"] = kind
        ap.add_argument(short, long_name, **kw)
    ap.add_argument("-v", "--version", action="version", version="merPCR version 1.0.0")
    ap.add_argument("--debug", action="store_true", help="Enable debug logging")
    ap.add_argument("--gpus", type=int, default=1,
                    help="NOT a reference flag: scan on this many GPUs of the node (one process per GPU, the genome "
                         "sharded by bp-balanced ranges with halos, final hit gather on rank 0); 0 = all visible")
    ap.add_argument("--true-strands", action="store_true",
                    help="NOT reference behaviour: report forward amplicons primer1 ... revcomp(primer2) as (+), "
                         "like NCBI me-PCR, instead of the reference's primer1 ... primer2")
    return ap


def main() -> int:
    """cli.py:217-266."""
    args = create_parser().parse_args(convert_mepcr_arguments(sys.argv[1:]))
    setup_logging(args.quiet, args.debug)
    logger = logging.getLogger("merpcr")
    from . import multi
    rank, world, local = multi.env_world()
    if world == 1 and args.gpus != 1:        # parent of a multi-GPU run: re-launch this command line as N ranks
        n = args.gpus
        if n == 0:
            import torch
            n = torch.cuda.device_count()
        if n > 1:
            return multi.launch(n, sys.argv[1:])
    try:
        if world > 1:                        # one rank of a multi-GPU run (launched above or by torchrun)
            multi.init_from_env()
            if rank != 0:
                logger.setLevel(logging.WARNING)
        mer_pcr = MerPCR(wordsize=args.wordsize, margin=args.margin, mismatches=args.mismatches,
                         three_prime_match=args.three_prime_match, iupac_mode=args.iupac,
                         default_pcr_size=args.default_pcr_size, threads=args.threads,
                         max_sts_line_length=args.max_sts_line_length, true_strands=args.true_strands,
                         **(dict(shard=(rank, world), device=local) if world > 1 else {}))
        if not mer_pcr.load_sts_file(args.sts_file):
            logger.error(f"Failed to load STS file: {args.sts_file}")
            return 1
        fasta_records = mer_pcr.load_fasta_file(args.fasta_file)
        if not fasta_records:
            logger.error(f"Failed to load FASTA file: {args.fasta_file}")
            return 1
        hit_count = mer_pcr.search(fasta_records, args.output)
        logger.info(f"Search complete: {hit_count} hits found")
        return 0
    except Exception as e:  # noqa: BLE001  (cli.py:260-266 maps every failure to exit 1)
        logger.error(f"Error: {str(e)}")
        if args.debug:
            import traceback
            traceback.print_exc()
        return 1
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
