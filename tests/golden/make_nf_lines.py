"""Extracts the literal `polycracker ...` / `bbmap.sh ...` command lines of the hot-path processes of the reference's
Nextflow script (polycracker/polycracker.nf:130, 161, 193, 200-201, 274, 318, 347) into tests/golden/nf_command_lines.txt,
one per line as `<nf line number>\t<command template>`.  They are the drop-in seam: tests/test_nextflow_seam.py formats
them with values parsed from a config file (a Python port of the script's findValue) and runs them through the CLI.
Run on the dev box (needs /root/reference)."""
import os
import re

NF = "/root/reference/polycracker/polycracker.nf"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nf_command_lines.txt")
STAGES = ("splitFasta", "writeKmerCount", "kmer2Fasta", "blast_kmers", "blast2bed", "generate_Kmer_Matrix")


def main():
    rows = []
    for no, line in enumerate(open(NF), 1):
        t = line.strip()
        if no > 360:
            break
        if t.startswith("echo "):
            continue
        if re.match(r"polycracker (%s) " % "|".join(STAGES), t) or ("bbmap.sh ref=" in t and no < 240):
            rows.append("%d\t%s" % (no, t))
    open(OUT, "w").write("\n".join(rows) + "\n")
    print(open(OUT).read())


if __name__ == "__main__":
    main()
