"""`kmer` command line (click), same commands and options as kmermaid/scripts/."""
import logging

logging.basicConfig(level=logging.INFO, format="%(message)s", datefmt="[%X]")
