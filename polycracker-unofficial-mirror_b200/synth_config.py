"""Shapes of the BASELINE.json workloads (pure constants: importable without loading the CUDA library, so the CPU
reference arm of bench.py reads the same numbers as the GPU arm)."""

#: defaults of SURVEY.md section 8d
DEFAULTS = dict(n_subgenomes=2, n_chrom=7, divergence=0.08, repeat_fraction=0.60, n_shared_families=200,
                n_specific_families=300, min_family_len=500, max_family_len=8000, max_copy_mutation=0.10,
                n_fraction=0.001, min_n_run=100, max_n_run=10000, softmask_fraction=0.10)

#: BASELINE.json configs -> generator + pipeline parameters (seed = 20260000 + cfg id)
CONFIGS = {
    "cfg1_algae_standin": dict(seed=20260001, genome_size=160_000_000, n_subgenomes=2, k=26, chunk_size=50000,
                               kmer_low_count=30),
    "cfg2_500Mbp_allotetraploid": dict(seed=20260002, genome_size=500_000_000, n_subgenomes=2, k=23,
                                       chunk_size=75000, kmer_low_count=30),
    "cfg3_4.5Gbp_tobacco": dict(seed=20260003, genome_size=4_500_000_000, n_subgenomes=2, k=23, chunk_size=75000,
                                kmer_low_count=100),
    "cfg4_15Gbp_wheat": dict(seed=20260004, genome_size=15_000_000_000, n_subgenomes=3, k=23, chunk_size=75000,
                             kmer_low_count=100),
}

#: field order of pc_synth_params (include/polycracker_b200.h)
PARAM_FIELDS = ["seed", "genome_size", "n_subgenomes", "n_chrom", "divergence", "repeat_fraction", "n_shared_families",
                "n_specific_families", "min_family_len", "max_family_len", "max_copy_mutation", "n_fraction",
                "min_n_run", "max_n_run", "softmask_fraction"]
