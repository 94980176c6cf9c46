"""`kmer batch` (kmermaid/scripts/kmer_batch.py:18-109): write sorted k-mer batch files."""
import logging
import os
import tempfile

import click

from kman_b200.batcher import BatcherThreading, FastaBatcher
from kman_b200.io import copy_batches, input_file_exists, set_tempdir
from kman_b200.scripts import arguments as args


@click.command(name="batch", context_settings=dict(help_option_names=["--help", "-h"]),
               help="Generate batches of k-mers from INPUT (FASTA files, one k-mer per record, sorted by sequence) in OUTPUT folder.")
@args.input_path()
@args.output_path(dir_okay=True)
@args.k()
@args.reverse()
@args.scan_mode()
@args.batch_size()
@args.batch_mode()
@args.threads()
@args.tmp()
@args.compress()
def run(input_path: str, output_path: str, k: int, reverse: bool = False, scan_mode: str = FastaBatcher.MODE.KMERS.name,
        batch_size: int = 1000000, batch_mode: str = BatcherThreading.FEED_MODE.APPEND.name, threads: int = 1,
        tmp: str = tempfile.gettempdir(), compress: bool = False) -> None:
    input_file_exists(input_path)
    set_tempdir(tmp)
    if os.path.isfile(output_path) or (os.path.isdir(output_path) and len(os.listdir(output_path)) > 0):
        raise AssertionError(f"output folder must be empty or non-existent: {output_path}")
    os.makedirs(output_path, exist_ok=True)
    batcher = FastaBatcher(scan_mode=FastaBatcher.MODE[scan_mode], reverse=reverse, size=batch_size, threads=threads, tmp=tmp)
    batcher.do(input_path, k, BatcherThreading.FEED_MODE[batch_mode])
    batches = [b for b in batcher.collection if b.current_size]
    for b in batches:
        b.write(True, True)
    copy_batches(batches, output_path, compress)
    logging.info("That's all! :smiley:")
