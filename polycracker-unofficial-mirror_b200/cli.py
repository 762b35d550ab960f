"""`polycracker` console script: the reference's click group (polycracker/polycracker.py:51-54) with the hot-path
sub-commands re-implemented on the GPU.  Option names, short flags and defaults are the reference's
(polycracker.py:113-122, 174-179, 212-218, 241-249, 255-262, 299-316) so polycracker.nf and the README recipes call
them unchanged.  `subgenomeExtraction` (scope row N1; options of polycracker.py:970-990) runs its count / compare /
map-back loop on the GPU too.  Sub-commands outside these rows (transform_plot, cluster, plots, ...) stay with the
reference package.
"""
import click

from . import stages, subgenome
from .binding import PolycrackerError

CONTEXT_SETTINGS = dict(help_option_names=["-h", "--help"], max_content_width=90)


@click.group(context_settings=CONTEXT_SETTINGS)
@click.version_option(version="1.1.3+b200")
def polycracker():
    pass


def _run(fn, **kw):
    try:
        return fn(**kw)
    except PolycrackerError as exc:          # the reference ignores tool failures; this path fails loudly
        raise click.ClickException(str(exc))


@polycracker.command(name="splitFasta")
@click.option("-i", "--fasta_file", help="Input polyploid fasta filename.", type=click.Path(exists=False))
@click.option("-d", "--fasta_path", default="./fasta_files/", help="Directory containing polyploids fasta file", show_default=True, type=click.Path(exists=False))
@click.option("-c", "--chunk_size", default=75000, help="Length of chunk of split fasta file", show_default=True)
@click.option("-p", "--prefilter", default=0, help="Remove chunks based on certain conditions as a preprocessing step to reduce memory", show_default=True)
@click.option("-m", "--min_chunk_size", default=0, help="If min_chunk_threshold and prefilter flag is on, use chunks with lengths greater than a certain size", show_default=True)
@click.option("-r", "--remove_non_chunk", default=0, help="If min_chunk_threshold and prefilter flag is on, remove chunks that are not the same size as chunk_size", show_default=True)
@click.option("-t", "--min_chunk_threshold", default=0, help="If prefilter is on, remove chunks with lengths less than a certain threshold", show_default=True)
@click.option("-l", "--low_memory", default=0, help="Must use with prefilter in order for prefilter to work.", show_default=True)
def splitFasta(fasta_file, fasta_path, chunk_size, prefilter, min_chunk_size, remove_non_chunk, min_chunk_threshold, low_memory):
    """Split fasta file into chunks of a specified length"""
    _run(stages.splitFasta, fasta_file=fasta_file, fasta_path=fasta_path, chunk_size=chunk_size, prefilter=prefilter,
         min_chunk_size=min_chunk_size, remove_non_chunk=remove_non_chunk, min_chunk_threshold=min_chunk_threshold,
         low_memory=low_memory)


@polycracker.command(name="writeKmerCount")
@click.option("-d", "--fasta_path", default="./fasta_files/", help="Directory containing chunked polyploids fasta file", show_default=True, type=click.Path(exists=False))
@click.option("-k", "--kmercount_path", default="./kmercount_files", help="Directory containing kmer count files", show_default=True, type=click.Path(exists=False))
@click.option("-l", "--kmer_length", default="23,31", help="Length of kmers to find; can include multiple lengths if comma delimited (e.g. 23,25,27)", show_default=True)
@click.option("-m", "--blast_mem", default="100", help="Accepted for compatibility (JVM heap of the BBTools run it replaces); ignored.", show_default=True)
def writeKmerCount(fasta_path, kmercount_path, kmer_length, blast_mem):
    """Counts the canonical k-mers of every chunked fasta file on the GPU (replaces kmercountexact.sh / jellyfish)"""
    _run(stages.writeKmerCount, fasta_path=fasta_path, kmercount_path=kmercount_path, kmer_length=kmer_length, blast_mem=blast_mem)


@polycracker.command(name="kmer2Fasta")
@click.option("-k", "--kmercount_path", default="./kmercount_files/", help="Directory containing kmer count files", show_default=True, type=click.Path(exists=False))
@click.option("-lc", "--kmer_low_count", default=100, help="Omit kmers from analysis that have less than x occurrences throughout genome.", show_default=True)
@click.option("-hb", "--high_bool", default=0, help="Enter 1 if you would like to use the kmer_high_count option.", show_default=True)
@click.option("-hc", "--kmer_high_count", default=2000000, help="If high_bool is set to one, omit kmers that have greater than x occurrences.", show_default=True)
@click.option("-s", "--sampling_sensitivity", default=1, help="If this option, x, is set greater than one, kmers are sampled at lower frequency, and total number of kmers included in the analysis is reduced to (total #)/(sampling_sensitivity), after threshold filtering", show_default=True)
def kmer2Fasta(kmercount_path, kmer_low_count, high_bool, kmer_high_count, sampling_sensitivity):
    """Converts kmer count file into a fasta file for blasting, after some threshold filtering and sampling"""
    _run(stages.kmer2Fasta, kmercount_path=kmercount_path, kmer_low_count=kmer_low_count, high_bool=high_bool,
         kmer_high_count=kmer_high_count, sampling_sensitivity=sampling_sensitivity)


@polycracker.command(name="blast_kmers")
@click.option("-m", "--blast_mem", default="48", help="Accepted for compatibility; ignored.", show_default=True)
@click.option("-r", "--reference_genome", help="Genome to blast against.", type=click.Path(exists=False))
@click.option("-k", "--kmer_fasta", help="Kmer fasta converted from kmer count file", type=click.Path(exists=False))
@click.option("-o", "--output_file", default="results.sam", show_default=True, help="Output sam file.", type=click.Path(exists=False))
@click.option("-kl", "--kmer_length", default=13, help="Accepted for compatibility (bbmap seed length); ignored.", show_default=True)
@click.option("-pm", "--perfect_mode", default=1, help="Perfect mode (must be 1: exact matching).", show_default=True)
@click.option("-t", "--threads", default=5, help="Accepted for compatibility; ignored.", show_default=True)
def blast_kmers(blast_mem, reference_genome, kmer_fasta, output_file, kmer_length, perfect_mode, threads):
    """Maps kmers fasta file to a reference genome (exact matches, either strand) on the GPU."""
    _run(stages.blast_kmers, blast_mem=blast_mem, reference_genome=reference_genome, kmer_fasta=kmer_fasta,
         output_file=output_file, kmer_length=kmer_length, perfect_mode=perfect_mode, threads=threads)


@polycracker.command(name="blast2bed")
@click.option("-i", "--blast_file", default="./blast_files/results.sam", help="Output file after blasting kmers to genome. Must be input into this command.", show_default=True, type=click.Path(exists=False))
@click.option("-b", "--bb", default=1, help="Whether bbtools were used in generating blasted sam file.", show_default=True)
@click.option("-l", "--low_memory", default=0, help="Do not merge the bed file. Takes longer to process but uses lower memory.", show_default=True)
@click.option("-o", "--output_bed_file", default="blasted.bed", show_default=True, help="Output bed file, not merged.", type=click.Path(exists=False))
@click.option("-om", "--output_merged_bed_file", default="blasted_merged.bed", show_default=True, help="Output bed file, merged.", type=click.Path(exists=False))
@click.option("-ex", "--external_call", is_flag=True, help="External call from non-polyCRACKER pipeline.")
def blast2bed(blast_file, bb, low_memory, output_bed_file, output_merged_bed_file, external_call):
    """Converts the blast results from blast or bbtools into a bed file"""
    _run(stages.blast2bed, blast_file=blast_file, bb=bb, low_memory=low_memory, output_bed_file=output_bed_file,
         output_merged_bed_file=output_merged_bed_file, external_call=external_call)


@polycracker.command(name="generate_Kmer_Matrix")
@click.option("-d", "--kmercount_path", default="./kmercount_files/", help="Directory containing kmer count files", type=click.Path(exists=False))
@click.option("-g", "--genome", help="File name of chunked genome fasta", type=click.Path(exists=False))
@click.option("-c", "--chunk_size", default=75000, help="Length of chunk of split/chunked fasta file", show_default=True)
@click.option("-mc", "--min_chunk_size", default=0, help="If min_chunk_threshold and prefilter flag is on, use chunks with lengths greater than a certain size", show_default=True)
@click.option("-r", "--remove_non_chunk", default=1, help="If min_chunk_threshold and prefilter flag is on, remove chunks that are not the same size as chunk_size", show_default=True)
@click.option("-t", "--min_chunk_threshold", default=0, help="If prefilter is on, remove chunks with lengths less than a certain threshold", show_default=True)
@click.option("-l", "--low_memory", default=0, help="Must use with prefilter in order for prefilter to work. Without prefilter, populates matrix using unmerged blasted bed (slow).", show_default=True)
@click.option("-p", "--prefilter", default=0, help="If selected, will now add chunks that were removed during preprocessing.", show_default=True)
@click.option("-f", "--fasta_path", default="./fasta_files/", help="Directory containing chunked polyploids fasta file", type=click.Path(exists=False))
@click.option("-gs", "--genome_split_name", help="File name of chunked genome fasta", type=click.Path(exists=False))
@click.option("-go", "--original_genome", help="File name of original, prechunked genome fasta", type=click.Path(exists=False))
@click.option("-i", "--input_bed_file", default="blasted.bed", show_default=True, help="Input bed file, not merged.", type=click.Path(exists=False))
@click.option("-im", "--input_merged_bed_file", default="blasted_merged.bed", show_default=True, help="Input bed file, merged.", type=click.Path(exists=False))
@click.option("-npz", "--output_sparse_matrix", default="clusteringMatrix.npz", show_default=True, help="Output sparse matrix file, in npz format.", type=click.Path(exists=False))
@click.option("-k", "--output_kmer_file", default="kmers.p", show_default=True, help="Output kmer pickle file.", type=click.Path(exists=False))
@click.option("-tf", "--tfidf", default=0, help="If set to one, no sequence length normalization will be done.", show_default=True)
def generate_Kmer_Matrix(kmercount_path, genome, chunk_size, min_chunk_size, remove_non_chunk, min_chunk_threshold, low_memory, prefilter, fasta_path, genome_split_name, original_genome, input_bed_file, input_merged_bed_file, output_sparse_matrix, output_kmer_file, tfidf):
    """From blasted bed file, generate the sparse k-mer (columns) x subsequence (rows) occurrence matrix"""
    _run(stages.generate_Kmer_Matrix, kmercount_path=kmercount_path, genome=genome, chunk_size=chunk_size,
         min_chunk_size=min_chunk_size, remove_non_chunk=remove_non_chunk, min_chunk_threshold=min_chunk_threshold,
         low_memory=low_memory, prefilter=prefilter, fasta_path=fasta_path, genome_split_name=genome_split_name,
         original_genome=original_genome, input_bed_file=input_bed_file, input_merged_bed_file=input_merged_bed_file,
         output_sparse_matrix=output_sparse_matrix, output_kmer_file=output_kmer_file, tfidf=tfidf)


@polycracker.command(name="subgenomeExtraction")
@click.option("-s", "--subgenome_folder", help="Subgenome folder where subgenome extraction is taking place. This should be the bootstrap_0 folder.", type=click.Path(exists=False))
@click.option("-os", "--original_subgenome_path", help="One level up from subgenomeFolder. Final outputs are stored here.", type=click.Path(exists=False))
@click.option("-p", "--fasta_path", default="./fasta_files/", help="Directory containing chunked polyploids fasta file", show_default=True, type=click.Path(exists=False))
@click.option("-g", "--genome_name", help="Filename of chunked polyploid fasta.", type=click.Path(exists=False))
@click.option("-go", "--original_genome", help="Filename of polyploid fasta; can be either original or chunked, just make sure original is set to 1 if using original genome.", type=click.Path(exists=False))
@click.option("-bb", "--bb", default=1, help="Whether bbtools were used in generating blasted sam file (must be 1).", show_default=True)
@click.option("-b", "--bootstrap", default=0, help="Number of times to bootstrap the subgenome extraction process.", show_default=True)
@click.option("-i", "--iteration", default=0, help="Current iteration of bootstrapping process. Please set to 0, when initially beginning.", show_default=True)
@click.option("-l", "--kmer_length", default="23,31", help="Length of kmers to find; can include multiple lengths if comma delimited (e.g. 23,25,27)", show_default=True)
@click.option("-r", "--run_final", default=0, help="Turn this to 0 when starting the command.", show_default=True)
@click.option("-o", "--original", default=0, help="Select 1 if trying to extract subgenomes based on original nonchunked genome.", show_default=True)
@click.option("-m", "--blast_mem", default="100", help="Accepted for compatibility; ignored.", show_default=True)
@click.option("-kl", "--kmer_low_count", default=100, help="Accepted for compatibility (unused by the reference here too).", show_default=True)
@click.option("-dk", "--diff_kmer_threshold", default=20, help="A kmer is differential if its count in a subgenome divided by its count in any other subgenome exceeds this value.", show_default=True)
@click.option("-u", "--unionbed_threshold", default="10,2", help="Two comma delimited values (absolute, ratio) used for binning regions into subgenomes.", show_default=True)
@click.option("-ds", "--diff_sample_rate", default=1, help="If greater than one, keep every x-th differential kmer after threshold filtering.", show_default=True)
@click.option("-kv", "--default_kmercount_value", default=3., help="Count used for a kmer that is not found in another subgenome's kmercount dictionary.", show_default=True)
@click.option("-sl", "--search_length", default=13, help="Accepted for compatibility (bbmap seed length); ignored.", show_default=True)
@click.option("-pm", "--perfect_mode", default=1, help="Perfect mode (must be 1: exact matching).", show_default=True)
def subgenomeExtraction(subgenome_folder, original_subgenome_path, fasta_path, genome_name, original_genome, bb, bootstrap, iteration, kmer_length, run_final, original, blast_mem, kmer_low_count, diff_kmer_threshold, unionbed_threshold, diff_sample_rate, default_kmercount_value, search_length, perfect_mode):
    """Extract subgenomes from genome, either chunked genome or original genome can be fed in as argument"""
    _run(subgenome.subgenomeExtraction, subgenome_folder=subgenome_folder, original_subgenome_path=original_subgenome_path,
         fasta_path=fasta_path, genome_name=genome_name, original_genome=original_genome, bb=bb, bootstrap=bootstrap,
         iteration=iteration, kmer_length=kmer_length, run_final=run_final, original=original, blast_mem=blast_mem,
         kmer_low_count=kmer_low_count, diff_kmer_threshold=diff_kmer_threshold, unionbed_threshold=unionbed_threshold,
         diff_sample_rate=diff_sample_rate, default_kmercount_value=default_kmercount_value,
         search_length=search_length, perfect_mode=perfect_mode)


@polycracker.command(name="number_repeatmers_per_subsequence")
@click.option("-m", "--merged_bed", default="blasted_merged.bed", help="Merged bed file containing bed subsequences and kmers, comma delimited.", show_default=True, type=click.Path(exists=False))
@click.option("-o", "--out_file", default="kmers_per_subsequence.png", help="Output histogram in png or pdf.", show_default=True, type=click.Path(exists=False))
@click.option("--kde", is_flag=True, help="Add kernel density estimation.")
def number_repeatmers_per_subsequence(merged_bed, out_file, kde):
    """Find histogram depicting number of repeat kmers per subsequence."""
    _run(stages.number_repeatmers_per_subsequence, merged_bed=merged_bed, out_file=out_file, kde=kde)


@polycracker.command(name="kcount_hist")
@click.option("-k", "--kcount_directory", help="Directory containing input kmer count files, named [kmer-length].kcount.", type=click.Path(exists=False))
@click.option("-o", "--out_file", default="kmer_histogram.png", help="Output histogram in png.", show_default=True, type=click.Path(exists=False))
@click.option("--log", is_flag=True, help="Scale x-axis to log base 10.")
@click.option("-r", "--kcount_range", default="3,10000", help="Comma delimited list of two items, specifying range of kmer counts.", show_default=True)
@click.option("-l", "--low_count", default=0., help="Input the kmer low count threshold and it will be plotted.", show_default=True)
@click.option("-kl", "--kmer_lengths", default="", help="Optional: comma delimited list of kmer lengths to include in analysis.")
@click.option("-s", "--sample_speed", default=1, help="Optional: increase sample speed to attempt to smooth kmer count histogram.", show_default=True)
def kcount_hist(kcount_directory, out_file, log, kcount_range, low_count, kmer_lengths, sample_speed):
    """Outputs a histogram plot of a given kmer count files."""
    _run(stages.kcount_hist, kcount_directory=kcount_directory, out_file=out_file, log=log, kcount_range=kcount_range,
         low_count=low_count, kmer_lengths=kmer_lengths, sample_speed=sample_speed)


@polycracker.command(name="hipmer_output_to_kcount")
@click.option("-i", "--hipmer_input", default="test.txt", help="Input file or directory from hipmer kmer counting run.", show_default=True, type=click.Path(exists=False))
@click.option("-o", "--kcount_output", default="test.final.kcount", help="Output kmer count file.", show_default=True, type=click.Path(exists=False))
@click.option("-d", "--run_on_dir", is_flag=True, help="Choose to run on all files in hipmer_input if you have specified a directory for the hipmer input. Directory can only contain hipmer files.")
def hipmer_output_to_kcount(hipmer_input, kcount_output, run_on_dir):
    """Converts hipmer kmer count output into a kmer count, kcount, file (polycracker/format.py:123-133)."""
    _run(stages.hipmer_output_to_kcount, hipmer_input=hipmer_input, kcount_output=kcount_output, run_on_dir=run_on_dir)


def main():
    polycracker()


if __name__ == "__main__":
    main()
