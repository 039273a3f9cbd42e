/* TEST / BENCH INFRASTRUCTURE — NOT PRODUCT CODE.
 * Times the reference CPU path for BASELINE.json configs[2] fairly: a pthreads loop over
 * the batch that calls LAPACKE_dgetrf + LAPACKE_dgetrs (function pointers resolved by the
 * caller in the OpenBLAS build Jalla would link, /root/reference/jalla.cabal:75) once per
 * 32x32 system, library threading off — the plan stated in BASELINE.md §3.            */
#include <pthread.h>
#include <stdlib.h>

typedef int (*getrf_fn)(int, int, int, double *, int, int *);
typedef int (*getrs_fn)(int, char, int, int, const double *, int, const int *, double *, int);

struct job { getrf_fn f; getrs_fn s; int n; double *A; int *piv; double *b; long lo, hi; int bad; };

static void *worker(void *p) {
  struct job *j = (struct job *)p;
  int n = j->n;
  for (long i = j->lo; i < j->hi; i++) {
    double *a = j->A + i * n * n;
    int *pv = j->piv + i * n;
    if (j->f(102, n, n, a, n, pv) != 0) j->bad++;
    if (j->s(102, 'N', n, 1, a, n, pv, j->b + i * n, n) != 0) j->bad++;
  }
  return 0;
}

int oracle_run_batched_lu(void *getrf, void *getrs, int n, double *A, int *piv, double *b, long batch, int threads) {
  if (threads < 1) threads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
  struct job *jobs = (struct job *)malloc(sizeof(struct job) * threads);
  for (int t = 0; t < threads; t++) {
    jobs[t].f = (getrf_fn)getrf; jobs[t].s = (getrs_fn)getrs; jobs[t].n = n; jobs[t].A = A; jobs[t].piv = piv; jobs[t].b = b;
    jobs[t].lo = batch * t / threads; jobs[t].hi = batch * (t + 1) / threads; jobs[t].bad = 0;
    pthread_create(&th[t], 0, worker, &jobs[t]);
  }
  int bad = 0;
  for (int t = 0; t < threads; t++) { pthread_join(th[t], 0); bad += jobs[t].bad; }
  free(th); free(jobs);
  return bad;
}

/* Same loop split over forked processes: OpenBLAS serialises concurrent LAPACK calls from
 * threads of one process on an internal buffer lock (measured here: no scaling from 1 to 8
 * threads), so the fair all-cores figure for the CPU path is process-parallel.          */
#include <sys/wait.h>
#include <unistd.h>
int oracle_run_batched_lu_procs(void *getrf, void *getrs, int n, double *A, int *piv, double *b, long batch, int procs) {
  if (procs < 1) procs = 1;
  pid_t *pids = (pid_t *)malloc(sizeof(pid_t) * procs);
  for (int t = 0; t < procs; t++) {
    pids[t] = fork();
    if (pids[t] == 0) {
      struct job j;
      j.f = (getrf_fn)getrf; j.s = (getrs_fn)getrs; j.n = n; j.A = A; j.piv = piv; j.b = b;
      j.lo = batch * t / procs; j.hi = batch * (t + 1) / procs; j.bad = 0;
      worker(&j);
      _exit(j.bad ? 1 : 0);
    }
  }
  int bad = 0;
  for (int t = 0; t < procs; t++) {
    int st = 0;
    if (pids[t] < 0 || waitpid(pids[t], &st, 0) < 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0) bad++;
  }
  free(pids);
  return bad;
}
