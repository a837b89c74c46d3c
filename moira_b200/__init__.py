"""moira_b200 -- B200-native (sm_100a CUDA) implementation of moira's per-read quality-filter
hot path, behind the reference's own interfaces.  See DESIGN.md and include/moira_b200.h."""
from .api import (Blocks, CollapseResult, FastqResult, RecordView, collapse_labels, fastq_headers, fastq_split, line_offsets, format_records, collapse, ContigParams, Context, FilterParams, FilterResult, MoiraError, PairResult,  # noqa: F401
                  PinnedBuffer, build_lut,
                  comm_init_all, comm_unique_id, pack_arrays, pack_q6, pack_reads, pack_sequences, parse_fasta_qual, parse_fastq,
                  reduce_counters_all, gz_deflate, gz_inflate, gz_scan)

__version__ = "0.1.0"
