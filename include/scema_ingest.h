/* scema_ingest.h — batch ingest of strain histories from files (host side of libscema_hist.so).
 *
 * Replaces, for a whole directory at once, the per-object text reader Strain6D::from_file
 * (reference headers/strain2spline.h:112-134) as the reference command lines drive it
 * (clustering/mpi_comparison_test.cc:67-88: readdir order, "strain_" name filter, ID = atoi of the
 * name without "strain_"; clustering/compare_all_histories.cc:41-57), and adds the converter from
 * the FE solver's own per-rank history log pr_<rank>.lhistory.csv (written by
 * FEProblem::output_lhistory, reference headers/FE_problem.h:1985-2045) to the same ragged batch.
 * A batch is exactly the argument list of scema_set_histories (include/scema_hist.h).
 *
 * Number syntax and rounding are those of `istream >> double` (the parser restates libstdc++'s
 * num_get grammar and yields strtod's correctly rounded value), reading of a file stops at the
 * first token that does not parse and a partially read line is dropped — as the reference's
 * `while (infile >> xx >> yy >> zz >> xy >> xz >> yz)`. No GPU is needed by these functions.
 * Return codes are those of scema_hist.h; scema_ingest_last_error() describes the last failure of
 * the calling thread.
 */
#ifndef SCEMA_INGEST_H
#define SCEMA_INGEST_H

#include <stdint.h>
#include "scema_hist.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct scema_batch scema_batch;

const char *scema_ingest_last_error(void);

/* Every "strain_*" entry of the directory in readdir order (the batch order of the reference
 * command lines). strain_directory is concatenated with the entry names as given, so it must end
 * in '/' (mpi_comparison_test.cc:81). n_threads = 0 uses all host cores. */
int scema_batch_read_dir(const char *strain_directory, uint32_t n_threads, scema_batch **out);
/* Explicit list of files; ids == NULL derives each ID from the file name as above. */
int scema_batch_read_files(const char *const *paths, const uint32_t *ids, uint64_t n, uint32_t n_threads,
                           scema_batch **out);
/* pr_<rank>.lhistory.csv files -> one history per qpid (ascending qpid; rows in file order, files
 * in argument order). column_prefix selects the tensor: "strain" (total strain = what
 * add_current_strain receives, FE_problem.h:1092-1098), "updstrain" or "stress"; NULL = "strain".
 * Components are reordered from the log's 00,01,02,11,12,22 to xx,yy,zz,xy,xz,yz. */
int scema_batch_from_lhistory(const char *const *csv_paths, uint64_t n_files, const char *column_prefix,
                              scema_batch **out);

uint64_t scema_batch_count(const scema_batch *b);       /* histories */
uint64_t scema_batch_total_steps(const scema_batch *b);  /* sum of lengths */
const double *scema_batch_steps(const scema_batch *b);   /* [total_steps][6] */
const uint64_t *scema_batch_offsets(const scema_batch *b); /* [count+1] */
const uint32_t *scema_batch_ids(const scema_batch *b);   /* [count] */
const char *scema_batch_name(const scema_batch *b, uint64_t i); /* file name / "qpid <id>" */

/* Write every history as "<out_directory>/strain_<ID>" in the format from_file reads (one line per
 * step, six values, 17 significant digits so that every double survives the round trip). */
int scema_batch_write_strain_files(const scema_batch *b, const char *out_directory);
/* scema_set_histories(ctx, steps, host, offsets, ids, count) of the batch. */
int scema_set_histories_from_batch(scema_ctx *ctx, const scema_batch *b);
void scema_batch_free(scema_batch *b);

#ifdef __cplusplus
}
#endif
#endif /* SCEMA_INGEST_H */
