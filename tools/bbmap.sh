#!/bin/bash
# Stand-in for BBTools' bbmap.sh on hosts that run polyCRACKER's mapping stage on the GPU (polycracker_b200).
#
# polycracker.nf calls `bbmap.sh ref=<genome> k=<k_search_length>` after kmer2Fasta (nf:200-201, 233-234) to build the
# on-disk index that the real `bbmap.sh ... perfectmode=t` of blast_kmers would read.  The GPU replacement of blast_kmers
# needs no index, so the index build is a no-op here.  Any OTHER use of bbmap.sh (an actual mapping run) is refused loudly:
# put the real BBTools first on PATH for that.
for arg in "$@"; do
    case "$arg" in
        in=*|in1=*|in2=*|out=*|outm=*) echo "bbmap.sh stand-in (polycracker_b200): only 'ref=... k=...' index builds are accepted; got '$arg'" >&2; exit 2 ;;
    esac
done
echo "bbmap.sh stand-in (polycracker_b200): index build skipped ($*)" >&2
exit 0
