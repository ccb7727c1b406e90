import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FIX = os.path.join(ROOT, "tests", "fixtures")
TF = os.path.join(FIX, "fasta_files", "Alp_V_ref.fasta")          # test/runtests.jl:8  tf
MINI_GENOME = os.path.join(FIX, "Alp_V_locus.fasta")               # test/runtests.jl:9  test_mini_genome
GENOME = os.path.join(FIX, "Loci.fasta")                           # test/runtests.jl:10 test_genome
EIGHT = os.path.join(FIX, "fasta_files", "8_ident_Alp_V_loci.fasta")  # test/runtests.jl:11 test_8_seqs

# test/runtests.jl:48
TEST_CONSENSUS = ("CAGGTGCAGCTGGTGGAGTCTGGGGGAGGCTTGGTGCAGCCTGGGGGGTCTCTGAGACTCTCCTGTGCAGCCTCTGGATTCACCTTCAGTAGC"
                  "TATGCCATGAGCTGGGTCCGCCAGGCTCCAGGGAAGGGGCTCGAGTGGGTCTCAGCTATTAATAGTGGTGGTGGTAGCACATACTATGCAGACT"
                  "CCGTGAAGGGCCGATTCACCATCTCCAGAGACAACGCCAAGAACACGCTGTATCTGCAAATGAACAGCCTGAAACCTGAGGGCACGGCCGTGTA"
                  "TTACTGTGGTAAAGAAGA")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the shared libraries are git-ignored build products: (re)build them when missing or older than their sources
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "kmergma.jl_b200", "csrc")], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
