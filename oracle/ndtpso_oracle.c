/* TEST INFRASTRUCTURE — not product code.
 *
 * Plain-C, flat-array, scalar fp64 restatement of libndtpso_slam's PSO scan-matching
 * hot path.  It is the CPU oracle the CUDA path is checked against; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (ndtpso_slam_b200/) never links, imports or calls it.
 *
 * Parity pin: the reference has no tests, golden vectors or fixtures of its own
 * (SURVEY.md section 4), so this file is pinned against the UNMODIFIED reference
 * sources compiled here (oracle/_ref, see oracle/Makefile): tests/test_oracle_vs_ref.py
 * requires bit-identical pose and cost on every case, and tests/golden/ holds the
 * reference's inputs/outputs as fixtures for machines without /root/reference.
 *
 * Each function cites the reference lines it follows (paths under /root/reference).
 * Compile with -ffp-contract=off: the reference is built for baseline x86-64
 * (CMakeLists.txt:5-9, no -march), so no a*b+c is ever fused there.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- *
 * glibc rand()/srand(): TYPE_3 additive-feedback generator
 * (glibc stdlib/random_r.c: __srandom_r, __random_r; degree 31, separation 3).
 * The reference draws every random number through Eigen's Random(), i.e.
 * std::rand() (core.cpp:14, core.cpp:84); nothing in the tree calls srand().
 * ------------------------------------------------------------------------- */
typedef struct {
  int32_t r[31];
  int f, b; /* front / rear indices into r */
} orc_rng;

void orc_srand(orc_rng* g, uint32_t seed) {
  if (seed == 0) seed = 1;
  int32_t word = (int32_t)seed; /* glibc keeps `word` in an int32_t: seeds >= 2^31 go negative */
  g->r[0] = word;
  for (int i = 1; i < 31; ++i) {
    const long long hi = word / 127773;
    const long long lo = word % 127773;
    long long t = 16807 * lo - 2836 * hi;
    if (t < 0) t += 2147483647;
    word = (int32_t)t;
    g->r[i] = word;
  }
  g->f = 3;
  g->b = 0;
  for (int i = 0; i < 310; ++i) {
    g->r[g->f] = (int32_t)((uint32_t)g->r[g->f] + (uint32_t)g->r[g->b]);
    g->f = (g->f + 1) % 31;
    g->b = (g->b + 1) % 31;
  }
}

int32_t orc_rand(orc_rng* g) {
  uint32_t v = (uint32_t)g->r[g->f] + (uint32_t)g->r[g->b];
  g->r[g->f] = (int32_t)v;
  g->f = (g->f + 1) % 31;
  g->b = (g->b + 1) % 31;
  return (int32_t)(v >> 1);
}

/* Fill out[0..n) with the rand() outputs that follow srand(seed). */
void orc_rand_stream(uint32_t seed, int32_t* out, int n) {
  orc_rng g;
  orc_srand(&g, seed);
  for (int i = 0; i < n; ++i) out[i] = orc_rand(&g);
}

/* ------------------------------------------------------------------------- *
 * Flat problem description: what cost_function reads from the two frames.
 * ------------------------------------------------------------------------- */
typedef struct {
  int32_t n_points;     /* points of the new frame, core.cpp:33-36 iteration order */
  int32_t w_cells;      /* NDTFrame::widthNumOfCells (ndtframe.cpp:27) */
  int32_t h_cells;      /* NDTFrame::heightNumOfCells (ndtframe.cpp:28) */
  int32_t _pad;
  double width_m;       /* NDTFrame::width  (uint16 metres) as double */
  double height_m;      /* NDTFrame::height (uint16 metres) as double */
  double cell_side;     /* NDTFrame::cell_side */
  double x_min, x_max;  /* NDTFrame::s_x_min/s_x_max (ndtframe.cpp:57-58) */
  double y_min, y_max;  /* NDTFrame::s_y_min/s_y_max (ndtframe.cpp:64-65) */
  const double* points; /* [n_points][2] */
  const double* mean;   /* [C][2]   NDTCell::mean */
  const double* inv_cov;/* [C][4]   NDTCell::s_inv_covar row-major 00,01,10,11 */
  const uint8_t* built; /* [C]      NDTCell::built */
} orc_problem;

typedef struct {
  int32_t iterations;   /* PSOConfig::iterations      config.h:28 */
  int32_t population;   /* PSOConfig::populationSize  config.h:29 */
  double w, c1, c2, w_dumping; /* PSOConfig::coeff    config.h:33-36 */
} orc_pso_config;

/* NDTFrame::getCellIndex, ndtframe.cpp:240-249 (strict bounds, double sum, then int). */
static int orc_cell_index(const orc_problem* p, double x, double y) {
  if ((x > p->x_min) && (x < p->x_max) && (y > p->y_min) && (y < p->y_max)) {
    return (int)(floor((x + (p->width_m / 2.)) / p->cell_side) +
                 (double)p->w_cells * (floor((y + (p->height_m / 2.)) / p->cell_side)));
  }
  return -1;
}

/* cost_function, core.cpp:26-48, with transform_point (core.h:28-31) and
 * NDTCell::normalDistribution (ndtcell.cpp:70-78) inlined.
 * An index at or past the end of the table is undefined behaviour in the
 * reference (reads past `cells`); here it is treated as "not built". */
double orc_cost(const orc_problem* p, const double* pose) {
  const double c = cos(pose[2]), s = sin(pose[2]);
  const int ncells = p->w_cells * p->h_cells;
  double trans_cost = 0.;
  for (int i = 0; i < p->n_points; ++i) {
    const double px = p->points[2 * i], py = p->points[2 * i + 1];
    const double x = px * c - py * s + pose[0];
    const double y = px * s + py * c + pose[1];
    const int idx = orc_cell_index(p, x, y);
    if (idx < 0 || idx >= ncells || !p->built[idx]) continue;
    const double d0 = x - p->mean[2 * idx], d1 = y - p->mean[2 * idx + 1];
    const double* S = p->inv_cov + 4 * idx;
    const double r0 = d0 * S[0] + d1 * S[2];
    const double r1 = d0 * S[1] + d1 * S[3];
    trans_cost -= exp(-(r0 * d0 + r1 * d1) / 2.);
  }
  return trans_cost;
}

void orc_cost_many(const orc_problem* p, const double* poses, int n, double* out) {
  for (int i = 0; i < n; ++i) out[i] = orc_cost(p, poses + 3 * i);
}

/* Eigen Random(): x + (y-x)*Scalar(rand())/Scalar(RAND_MAX), x=-1, y=1. */
static double orc_unit(int32_t r) { return -1.0 + 2.0 * (double)r / 2147483647.0; }

typedef struct {
  int32_t gbest_updates;   /* gbest improvements inside the iteration loop */
  int32_t pbest_updates;
  int32_t rand_draws;
  int32_t _pad;
} orc_stats;

/* pso_optimization, core.cpp:50-116 (single-thread order), Particle ctor core.cpp:13-23.
 * Random numbers come from `stream` (rand() outputs, consumed in order) when it is
 * non-NULL, else from the TYPE_3 generator seeded with `seed`.
 * out_pose = global_best.best_position, out_cost = global_best.best_cost. */
int orc_pso(const orc_problem* p, const double* guess, const double* deviation, const orc_pso_config* cf, uint32_t seed,
            const int32_t* stream, double* out_pose, double* out_cost, orc_stats* stats) {
  const int P = cf->population, I = cf->iterations;
  orc_rng g;
  int drawn = 0;
  if (!stream) orc_srand(&g, seed);
#define NEXT_RAND() (stream ? stream[drawn++] : (drawn++, orc_rand(&g)))

  double* x = (double*)malloc(sizeof(double) * 3 * (size_t)(P > 0 ? P : 1));
  double* v = (double*)malloc(sizeof(double) * 3 * (size_t)(P > 0 ? P : 1));
  double* pb = (double*)malloc(sizeof(double) * 3 * (size_t)(P > 0 ? P : 1));
  double* pbc = (double*)malloc(sizeof(double) * (size_t)(P > 0 ? P : 1));
  if (!x || !v || !pb || !pbc) return -1;

  const double zero_devi[3] = {1E-4, 1E-4, 1E-5}; /* core.cpp:53 */
  double gb[3], gbc;
  for (int k = 0; k < 3; ++k) gb[k] = guess[k] + orc_unit(NEXT_RAND()) * zero_devi[k]; /* core.cpp:58 */
  gbc = orc_cost(p, gb);

  for (int j = 0; j < P; ++j) { /* core.cpp:60-69 */
    for (int k = 0; k < 3; ++k) {
      x[3 * j + k] = guess[k] + orc_unit(NEXT_RAND()) * deviation[k];
      v[3 * j + k] = 0.;
      pb[3 * j + k] = x[3 * j + k];
    }
    pbc[j] = orc_cost(p, x + 3 * j);
    if (pbc[j] < gbc) {
      gbc = pbc[j];
      memcpy(gb, pb + 3 * j, sizeof gb);
    }
  }

  int n_gb = 0, n_pb = 0;
  double w = cf->w;
  for (int it = 0; it < I; ++it) { /* core.cpp:78-109 */
    for (int j = 0; j < P; ++j) {
      for (int k = 0; k < 3; ++k) {
        const double rx = fabs(orc_unit(NEXT_RAND()));
        const double ry = fabs(orc_unit(NEXT_RAND()));
        v[3 * j + k] = w * v[3 * j + k] + cf->c1 * rx * (pb[3 * j + k] - x[3 * j + k]) + cf->c2 * ry * (gb[k] - x[3 * j + k]);
        x[3 * j + k] = x[3 * j + k] + v[3 * j + k];
      }
      const double cost = orc_cost(p, x + 3 * j);
      if (cost < pbc[j]) {
        pbc[j] = cost;
        memcpy(pb + 3 * j, x + 3 * j, 3 * sizeof(double));
        ++n_pb;
        if (cost < gbc) {
          gbc = cost;
          memcpy(gb, x + 3 * j, sizeof gb);
          ++n_gb;
        }
      }
    }
    w *= cf->w_dumping;
  }
#undef NEXT_RAND

  memcpy(out_pose, gb, sizeof gb);
  if (out_cost) *out_cost = gbc;
  if (stats) {
    stats->gbest_updates = n_gb;
    stats->pbest_updates = n_pb;
    stats->rand_draws = drawn;
    stats->_pad = 0;
  }
  free(x);
  free(v);
  free(pb);
  free(pbc);
  return 0;
}

/* glir_pso_optimization, core.cpp:118-186 (the reference marks it "UNTESTED" and never calls it), single-thread order.
 * The reference hard-wires the population to PSO_POPULATION_SIZE (30, config.h:21); here it is a parameter.
 *   :125      global_best is a Particle of its own, drawn with the CALLER's deviation (zero_devi, :122-123, is never used)
 *   :132-141  P + 1 particles are constructed (3 draws and one cost evaluation each: 3(P + 2) draws in all); the one built in
 *             pass i is particles[i + 1], the one compared with global_best is particles[i]: the last particle never
 *             takes part, neither here nor in the iterations (j < P, :145)
 *   :146      omega = 1.1 - gbest_cost / (pbest_average_j / (j + 1)), j + 1 an unsigned converted to double
 *   :147      c1 = c2 = 1.0 + gbest_cost / pbest_cost_j
 *   :148-157  per coordinate: two draws (Array2d::Random().abs(): x then y), best_ratio = pbest_j[k] / gbest[k],
 *             v = omega*v + c1*rx*(best_ratio*pbest - x) + c2*ry*((1./best_ratio)*gbest - x); x += v
 *   :161-174  pbest, pbest_average += pbest_cost, gbest (strict <, taken from the particle's best fields)
 * out_cost = global_best.best_cost. */
int orc_glir(const orc_problem* p, const double* guess, const double* deviation, int population, int iterations, uint32_t seed,
             const int32_t* stream, double* out_pose, double* out_cost, orc_stats* stats) {
  const int P = population, I = iterations;
  orc_rng g;
  int drawn = 0;
  if (!stream) orc_srand(&g, seed);
#define NEXT_RAND() (stream ? stream[drawn++] : (drawn++, orc_rand(&g)))
  const size_t n = (size_t)P + 1;
  double* x = (double*)malloc(sizeof(double) * 3 * n);
  double* v = (double*)malloc(sizeof(double) * 3 * n);
  double* pb = (double*)malloc(sizeof(double) * 3 * n);
  double* pbc = (double*)malloc(sizeof(double) * n);
  double* cst = (double*)malloc(sizeof(double) * n);
  double* pavg = (double*)malloc(sizeof(double) * n);
  if (!x || !v || !pb || !pbc || !cst || !pavg) return -1;

  double gb[3], gbc;
  for (int k = 0; k < 3; ++k) gb[k] = guess[k] + orc_unit(NEXT_RAND()) * deviation[k]; /* :125 */
  gbc = orc_cost(p, gb);
#define NEW_PARTICLE(j)                                                                \
  do {                                                                                 \
    for (int k = 0; k < 3; ++k) {                                                      \
      x[3 * (j) + k] = guess[k] + orc_unit(NEXT_RAND()) * deviation[k];                \
      v[3 * (j) + k] = 0.;                                                             \
      pb[3 * (j) + k] = x[3 * (j) + k];                                                \
    }                                                                                  \
    cst[j] = pbc[j] = pavg[j] = orc_cost(p, x + 3 * (j)); /* core.cpp:18-22 */         \
  } while (0)
  NEW_PARTICLE(0); /* :132 */
  for (int i = 0; i < P; ++i) { /* :134-141 */
    NEW_PARTICLE(i + 1);
    if (cst[i] < gbc) {
      gbc = pbc[i];
      memcpy(gb, pb + 3 * i, sizeof gb);
    }
  }
#undef NEW_PARTICLE

  int n_gb = 0, n_pb = 0;
  for (int it = 0; it < I; ++it) { /* :143-177 */
    for (int j = 0; j < P; ++j) {
      const double omega = 1.1 - gbc / (pavg[j] / (double)(unsigned)(j + 1));
      const double c12 = 1.0 + gbc / pbc[j];
      for (int k = 0; k < 3; ++k) {
        const double rx = fabs(orc_unit(NEXT_RAND()));
        const double ry = fabs(orc_unit(NEXT_RAND()));
        const double best_ratio = pb[3 * j + k] / gb[k];
        v[3 * j + k] = omega * v[3 * j + k] + c12 * rx * (best_ratio * pb[3 * j + k] - x[3 * j + k]) +
                       c12 * ry * ((1. / best_ratio) * gb[k] - x[3 * j + k]);
        x[3 * j + k] = x[3 * j + k] + v[3 * j + k];
      }
      cst[j] = orc_cost(p, x + 3 * j);
      if (cst[j] < pbc[j]) {
        pbc[j] = cst[j];
        memcpy(pb + 3 * j, x + 3 * j, 3 * sizeof(double));
        ++n_pb;
      }
      pavg[j] += pbc[j];
      if (cst[j] < gbc) {
        gbc = pbc[j];
        memcpy(gb, pb + 3 * j, sizeof gb);
        ++n_gb;
      }
    }
  }
#undef NEXT_RAND

  memcpy(out_pose, gb, sizeof gb);
  if (out_cost) *out_cost = gbc;
  if (stats) {
    stats->gbest_updates = n_gb;
    stats->pbest_updates = n_pb;
    stats->rand_draws = drawn;
    stats->_pad = 0;
  }
  free(x);
  free(v);
  free(pb);
  free(pbc);
  free(cst);
  free(pavg);
  return 0;
}

/* NDTFrame::align, ndtframe.cpp:251-266: the deviation rule and the s_* bookkeeping
 * around pso_optimization (which it calls with the DEFAULT PSOConfig: 30 x 50,
 * w=.8, c1=c2=2, w_dumping=1; config.h:20-37).  state = {s_iter, s_prev_pose[3], s_pose_diff[3]}. */
typedef struct {
  int32_t s_iter;
  int32_t _pad;
  double s_prev_pose[3];
  double s_pose_diff[3];
} orc_align_state;

void orc_align_deviation(const orc_align_state* st, double* dev) {
  if (st->s_iter < 2) {
    dev[0] = .1;
    dev[1] = .1;
    dev[2] = 3.1415E-3;
  } else {
    for (int k = 0; k < 3; ++k) dev[k] = fabs(st->s_pose_diff[k] * 2.);
  }
}

int orc_align(orc_align_state* st, const orc_problem* p, const double* guess, uint32_t seed, const int32_t* stream, double* out_pose) {
  double dev[3];
  orc_align_deviation(st, dev);
  ++st->s_iter;
  orc_pso_config cf = {50, 30, .8, 2., 2., 1.};
  int rc = orc_pso(p, guess, dev, &cf, seed, stream, out_pose, 0, 0);
  for (int k = 0; k < 3; ++k) {
    st->s_pose_diff[k] = out_pose[k] - st->s_prev_pose[k];
    st->s_prev_pose[k] = out_pose[k];
  }
  return rc;
}
