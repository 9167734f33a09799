#!/usr/bin/env python3
"""Recipe that compiles the UNMODIFIED reference (cychan-lbnl/devastator) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may run it.

What it does (world=threads backend, define set of SURVEY.md Appendix C = what the
reference's own `brutal` tool would have generated, brutal.py:43-166):

  * compiles the eight runtime translation units where they lie under /root/reference
    (pdes gvt threads diagnostic datarow opnew world_threads reduce) once per rank count R
    into oracle/_ref/obj_R<R>/;
  * links the reference's own apps, sources untouched:
        test/phold.cxx   -> oracle/_ref/ref_phold_test_R<R>
        bench/phold.cxx  -> oracle/_ref/ref_phold_bench_R<R>
  * PHOLD-T: the reference's test/phold.cxx MODEL with its four
    constants, root time and tiebreak made compile-time parameters.  The model text is never
    copied into this repo: it is read from /root/reference/test/phold.cxx at build time, the
    constants are substituted by the regexes below, the reference's main() is cut off and the
    driver in PHOLD_T_MAIN (ours: one plain drain, timing, per-LP state dump) is appended.  The
    patched file only ever exists under oracle/_ref/gen/ (git-ignored).
        -> oracle/_ref/ref_phold_t_<tag>_R<R>

Outputs only under oracle/_ref/ (git-ignored, NOT gpurun-ignored: the binaries travel to the
GPU box, /root/reference does not exist there).

Usage:  python oracle/build_ref.py [--ranks 2 8] [--only phold_test,phold_t,...] [--force]
"""
import argparse
import hashlib
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("DEVA_REFERENCE_ROOT", "/root/reference")

# Appendix C define set (brutal.py:43-166 defaults).
DEFINES = [
    "-DDEBUG=0", "-DNDEBUG=1", "-DDEVA_OPNEW_DEVA=1", "-DDEVA_DUMMY_EXEC=0", "-DDRAIN_TIMER=0",
    "-DTIMELINE=0", "-DDEVA_WORLD=1", "-DDEVA_WORLD_THREADS=1", "-DDEVA_THREADS_SPSC=1",
    "-DDEVA_THREADS_ALLOC_OPNEW_ASYM=1",
    "-DDEVA_THREADS_SPSC_PREFETCH=0", "-DDEVA_THREADS_SPSC_ORDER_DFS=1",
    "-DDEVA_THREADS_SIGNAL_REAP_MEMCPY=1", '-DDEVA_GIT_VERSION="ref"',
]
# -O3 builds of the reference use LTO (brutal.py:62,82)
CXXFLAGS = ["-std=c++14", "-O3", "-flto", "-pthread", "-Wno-aligned-new", "-w"]


def spsc_bits(R):
    """tsigbits as the reference's build tool derives it from the thread count (brutal.py:118-121):
    64*8/threads, rounded down to one of 8 / 16 / 32."""
    t = 64 * 8 / R
    return (8, 16, 32)[sum(x < t for x in (8, 16))]
RUNTIME_TUS = [
    "src/devastator/pdes.cxx", "src/devastator/gvt.cxx", "src/devastator/threads.cxx",
    "src/devastator/diagnostic.cxx", "src/devastator/datarow.cxx", "src/devastator/opnew.cxx",
    "src/devastator/world/world_threads.cxx", "src/devastator/world/reduce.cxx",
]

# PHOLD-T configurations: tag -> parameters.  `sub`=1 enables the reference's documented
# `uint64_t subtime() const { return ray; }` tiebreak (test/phold.cxx:74-76); `rootdiv`=1 makes
# the root time ray/A instead of ray so that all K roots of an LP start in [0,K).
PHOLD_T_CONFIGS = {
    # small cases the C restatement and the GPU parity tests are checked against
    "tiny":   dict(A=64,      K=4,  P=0.5,  LAMBDA=10,  END=400,  sub=1, rootdiv=1),
    "small":  dict(A=1000,    K=16, P=0.10, LAMBDA=100, END=300,  sub=1, rootdiv=1),
    "smallq": dict(A=1000,    K=16, P=0.50, LAMBDA=100, END=300,  sub=1, rootdiv=1),
    "seqid":  dict(A=1000,    K=2,  P=0.5,  LAMBDA=100, END=10000, sub=0, rootdiv=0),  # == test/phold.cxx
    # more parameter regimes, only to pin the oracle against the reference (tests/test_oracle.py)
    "allrem": dict(A=257,     K=3,  P=1.0,  LAMBDA=50,  END=500,  sub=1, rootdiv=1),   # every child leaves its LP, odd A
    "norem":  dict(A=300,     K=5,  P=0.0,  LAMBDA=20,  END=400,  sub=1, rootdiv=1),   # no child ever does
    "dense":  dict(A=500,     K=8,  P=0.3,  LAMBDA=5,   END=150,  sub=1, rootdiv=1),   # short delays: many same-time events
    "seq3":   dict(A=777,     K=3,  P=0.25, LAMBDA=100, END=3000, sub=0, rootdiv=0),   # sequence-id tiebreak, A % ranks != 0
    "rayt":   dict(A=2000,    K=1,  P=0.5,  LAMBDA=200, END=2000, sub=1, rootdiv=0),   # root time = ray id
    "wide":   dict(A=20000,   K=8,  P=0.5,  LAMBDA=100, END=200,  sub=1, rootdiv=1),
    # SURVEY Appendix B golden: commits 6 241 953, checksum 3349230152040954854
    "mid":    dict(A=100000,  K=16, P=0.10, LAMBDA=100, END=300,  sub=1, rootdiv=1),
    # BASELINE config 2 (1M LPs, 16 ev/LP, 10% remote); END is a run-time env knob (PHOLD_T_END)
    "cfg2":   dict(A=1048576, K=16, P=0.10, LAMBDA=100, END=1000, sub=1, rootdiv=1),
    # BASELINE config 3 (8M LPs, 50% remote): the CPU baseline of the multi-GPU bench lines
    "cfg3":   dict(A=8388608, K=16, P=0.50, LAMBDA=100, END=1000, sub=1, rootdiv=1),
}
BASELINE_TAGS = ("cfg2", "cfg3")   # built for the host core counts a GPU box may have, not for the small rank counts

PHOLD_T_MAIN = r'''
// ---- driver appended by oracle/build_ref.py (not reference text) --------------------------------
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
static uint64_t *g_dump = nullptr;           // [actor_n][3] = rng.a, rng.b, check
static uint64_t g_end_time_rt = 0;
int main() {
  const char *dump_path = std::getenv("PHOLD_T_DUMP");
  const char *drain_end_s = std::getenv("PHOLD_T_DRAIN_END");
  uint64_t drain_end = drain_end_s ? std::strtoull(drain_end_s, nullptr, 10) : ~uint64_t(0);
  g_dump = (uint64_t*)std::calloc(size_t(actor_n)*3, sizeof(uint64_t));
  int seg_n = 0; uint64_t seg_dt = 0, seg_t0 = 0;
  if(const char *sg = std::getenv("PHOLD_T_SEGMENTS")) {   // "n:dt[:t0]"  n slices of dt after an untimed lead-in [0,t0)
    seg_n = std::atoi(sg);
    const char *c = std::strchr(sg, ':');
    seg_dt = c ? std::strtoull(c+1, nullptr, 10) : 1;
    const char *c2 = c ? std::strchr(c+1, ':') : nullptr;
    seg_t0 = c2 ? std::strtoull(c2+1, nullptr, 10) : 0;
  }
  auto doit = [&]() {
    pdes::chitter_secs = -1;
    pdes::init(actor_per_rank);
    int actor_lb = rank_me()*actor_per_rank;
    int actor_ub = std::min(actor_n, (rank_me()+1)*actor_per_rank);
    for(int actor=actor_lb; actor < actor_ub; actor++) {
      int cd = actor - actor_lb;
      state_cur[cd] = rng_state{actor};
      check[cd] = actor;
      pdes::register_state(cd, &state_cur[cd]);
      pdes::register_state(cd, &check[cd]);
    }
    for(long ray=0; ray < ray_n; ray++) {
      int actor = int(ray % actor_n);
      if(actor_lb <= actor && actor < actor_ub)
        pdes::root_event(actor - actor_lb, PHOLD_ROOT_TIME(ray), event{int(ray), actor});
    }
    deva::barrier();
    uint64_t gvt = 0;
    double secs = 0;
    if(seg_n > 0) {
      // segment mode (bench.py reference arm): setup once, then seg_n timed drain(t_end) slices of
      // seg_dt simulated-time units each; one JSON line per slice
      uint64_t commits_prev = 0;
      if(seg_t0 > 0) {   // untimed lead-in: skips the start-up transient (all roots sit in [0,K))
        pdes::drain(seg_t0, false);
        commits_prev = deva::reduce_sum(pdes::local_stats()).committed_n;
      }
      for(int sgi=0; sgi < seg_n; sgi++) {
        uint64_t t_end_i = seg_t0 + uint64_t(sgi+1)*seg_dt;
        deva::barrier();
        auto t0 = std::chrono::steady_clock::now();
        gvt = pdes::drain(t_end_i, false);
        double s1 = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        s1 = deva::reduce_max(s1);
        pdes::statistics st1 = deva::reduce_sum(pdes::local_stats());
        if(deva::rank_me() == 0) {
          std::printf("{\"segment\": %d, \"t_end\": %llu, \"commits\": %llu, \"drain_secs\": %.6f, \"ranks\": %d}\n",
                      sgi, (unsigned long long)t_end_i, (unsigned long long)(st1.committed_n - commits_prev), s1, rank_n);
          std::fflush(stdout);
        }
        commits_prev = st1.committed_n;
        secs += s1;
      }
    }
    else {
      auto t0 = std::chrono::steady_clock::now();
      gvt = pdes::drain(drain_end, false);
      secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      secs = deva::reduce_max(secs);
    }
    pdes::finalize();
    pdes::statistics stats = deva::reduce_sum(pdes::local_stats());
    uint64_t chk = checksum();
    for(int actor=actor_lb; actor < actor_ub; actor++) {
      int cd = actor - actor_lb;
      g_dump[3*size_t(actor)+0] = state_cur[cd].a;
      g_dump[3*size_t(actor)+1] = state_cur[cd].b;
      g_dump[3*size_t(actor)+2] = check[cd];
    }
    deva::barrier();
    if(deva::rank_me() == 0) {
      std::printf("{\"ranks\": %d, \"actor_n\": %d, \"ray_n\": %ld, \"p_remote\": %.17g, \"lambda\": %.17g, "
                  "\"end_time\": %llu, \"drain_end\": %llu, \"user_subtime\": %d, "
                  "\"executed\": %llu, \"commits\": %llu, \"deterministic\": %d, "
                  "\"checksum\": %llu, \"gvt\": %llu, \"drain_secs\": %.6f}\n",
                  rank_n, actor_n, (long)ray_n, (double)percent_remote, (double)lambda,
                  (unsigned long long)end_time, (unsigned long long)drain_end, PHOLD_USER_SUBTIME,
                  (unsigned long long)stats.executed_n, (unsigned long long)stats.committed_n,
                  int(stats.deterministic), (unsigned long long)chk, (unsigned long long)gvt, secs);
      if(dump_path) {
        FILE *f = std::fopen(dump_path, "wb");
        std::fwrite(g_dump, sizeof(uint64_t), size_t(actor_n)*3, f);
        std::fclose(f);
      }
    }
  };
  deva::run(doit);
  return 0;
}
'''


def sh(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout)
        raise SystemExit("oracle/build_ref.py: compile failed")
    return r.stdout


def newer(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs if os.path.exists(s)) or \
        os.path.getmtime(__file__) > t


def build_runtime(R, force):
    objdir = os.path.join(OUT, f"obj_R{R}")
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    objs = []
    for tu in RUNTIME_TUS:
        src = os.path.join(REF, tu)
        obj = os.path.join(objdir, os.path.basename(tu).replace(".cxx", ".o"))
        objs.append(obj)
        if force or newer(obj, [src]):
            jobs.append(["g++", *CXXFLAGS, *DEFINES, f"-DDEVA_THREAD_N={R}", f"-DDEVA_THREADS_SPSC_BITS={spsc_bits(R)}",
                         f"-I{REF}/src", "-c", src, "-o", obj])
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(sh, jobs))
    return objs


def link_app(R, objs, src, exe, extra=(), force=False):
    if not (force or newer(exe, [src] + objs)):
        return exe
    sh(["g++", *CXXFLAGS, *DEFINES, f"-DDEVA_THREAD_N={R}", f"-DDEVA_THREADS_SPSC_BITS={spsc_bits(R)}", f"-I{REF}/src",
        f"-I{REF}/bench", *extra, src, *objs, "-o", exe])
    return exe


def gen_phold_t():
    """Patch the reference's PHOLD model into a parametrised translation unit (never committed)."""
    text = open(os.path.join(REF, "test/phold.cxx")).read()
    cut = text.index("int main()")
    model = text[:cut]

    def sub1(pat, rep, s):
        s2, n = re.subn(pat, rep, s, count=1, flags=re.M)
        if n != 1:
            raise SystemExit(f"oracle/build_ref.py: pattern not found in reference test/phold.cxx: {pat}")
        return s2
    model = sub1(r"constexpr int actor_n = \d+;", "constexpr int actor_n = PHOLD_A;", model)
    model = sub1(r"constexpr int ray_n = 2\*actor_n;", "constexpr long ray_n = long(PHOLD_K)*actor_n;", model)
    model = sub1(r"constexpr double percent_remote = [\d.]+;", "constexpr double percent_remote = PHOLD_P;", model)
    model = sub1(r"constexpr double lambda = \d+;", "constexpr double lambda = PHOLD_LAMBDA;", model)
    model = sub1(r"constexpr uint64_t end_time = uint64_t\(100\*lambda\);",
                 "static uint64_t end_time = std::getenv(\"PHOLD_T_END\") ? "
                 "std::strtoull(std::getenv(\"PHOLD_T_END\"), nullptr, 10) : uint64_t(PHOLD_END);", model)
    # the reference's own documented alternative tiebreak, test/phold.cxx:75
    model = sub1(r"^\s*//uint64_t subtime\(\) const \{ return ray; \}",
                 "#if PHOLD_USER_SUBTIME\n  uint64_t subtime() const { return ray; }\n#endif", model)
    os.makedirs(os.path.join(OUT, "gen"), exist_ok=True)
    path = os.path.join(OUT, "gen", "phold_t.cxx")
    body = "#include <cstdlib>\n" + model + PHOLD_T_MAIN
    if not os.path.exists(path) or open(path).read() != body:
        open(path, "w").write(body)
    return path


PHOLD_B_MAIN = r'''
// ---- driver appended by oracle/build_ref.py (not reference text) --------------------------------
// PHOLD-B: bench/phold.cxx's model made reproducible - the wall-clock cut-off is replaced by an end
// time (an event sends its child iff time + dt < PHOLD_B_END; the rng is consumed either way), the
// report by a JSON line + a dump of every LP's rng state.  Setup as in the reference's main().
#include <cstdio>
#include <cstdlib>
int main() {
  const char *dump_path = std::getenv("PHOLD_B_DUMP");
  const char *drain_end_s = std::getenv("PHOLD_B_DRAIN_END");
  const uint64_t drain_end = drain_end_s ? std::strtoull(drain_end_s, nullptr, 10) : ~uint64_t(0);
  g_phold_b_end = std::getenv("PHOLD_B_END") ? std::strtoull(std::getenv("PHOLD_B_END"), nullptr, 10) : ~uint64_t(0);
  static uint64_t *dump = nullptr;
  auto doit = [&]() {
    if(deva::rank_me_local() == 0) {
      lp_per_rank = deva::os_env<int>("lp_per_rank", 1000);
      ray_per_lp = deva::os_env<int>("ray_per_lp", 2);
      peer_stddev = deva::os_env<double>("peer_stddev", 2.0);
      dump = (uint64_t*)std::calloc(size_t(lp_per_rank)*rank_n*2, sizeof(uint64_t));
    }
    deva::barrier();
    pdes::chitter_secs = -1;
    pdes::init(lp_per_rank);
    state_cur.reset(new rng_state[lp_per_rank]);
    for(int cd=0; cd < lp_per_rank; cd++) {
      int lp = rank_me()*lp_per_rank + cd;
      state_cur[cd] = rng_state{lp};
      pdes::register_state(cd, &state_cur[cd]);
      for(int lp_ray=0; lp_ray < ray_per_lp; lp_ray++) {
        int ray = lp*ray_per_lp + lp_ray;
        pdes::root_event(cd, ray, bounce{ray, lp});
      }
    }
    deva::barrier();
    uint64_t gvt = pdes::drain(drain_end, false);
    pdes::finalize();
    pdes::statistics stats = deva::reduce_sum(pdes::local_stats());
    for(int cd=0; cd < lp_per_rank; cd++) {
      size_t lp = size_t(rank_me())*lp_per_rank + cd;
      dump[2*lp+0] = state_cur[cd].a;
      dump[2*lp+1] = state_cur[cd].b;
    }
    deva::barrier();
    if(deva::rank_me() == 0) {
      std::printf("{\"ranks\": %d, \"lp_per_rank\": %d, \"ray_per_lp\": %d, \"peer_stddev\": %.17g, "
                  "\"end_time\": %llu, \"drain_end\": %llu, \"commits\": %llu, \"deterministic\": %d, \"gvt\": %llu}\n",
                  rank_n, lp_per_rank, ray_per_lp, peer_stddev, (unsigned long long)g_phold_b_end,
                  (unsigned long long)drain_end, (unsigned long long)stats.committed_n, int(stats.deterministic),
                  (unsigned long long)gvt);
      if(dump_path) {
        FILE *f = std::fopen(dump_path, "wb");
        std::fwrite(dump, sizeof(uint64_t), size_t(lp_per_rank)*rank_n*2, f);
        std::fclose(f);
      }
    }
  };
  deva::run(doit);
  return 0;
}
'''


def gen_phold_b():
    """bench/phold.cxx's model with an end time instead of the wall-clock cut-off (never committed)."""
    text = open(os.path.join(REF, "bench/phold.cxx")).read()
    model = text[:text.index("int main()")]

    def sub1(pat, rep, s, flags=re.M):
        s2, n = re.subn(pat, rep, s, count=1, flags=flags)
        if n != 1:
            raise SystemExit(f"oracle/build_ref.py: pattern not found in reference bench/phold.cxx: {pat}")
        return s2
    # the cut-off block (bench/phold.cxx:95-104) goes; the send becomes conditional on the end time
    model = sub1(r"static __thread int skips = 0;.*?\n      \}\n    \}\n", "", model, flags=re.S)
    model = sub1(r"^(\s*)cxt\.send\(", r"\1if(cxt.time + dt < g_phold_b_end) cxt.send(", model)
    model = sub1(r"^struct bounce \{", "static uint64_t g_phold_b_end = ~uint64_t(0);\nstruct bounce {", model)
    os.makedirs(os.path.join(OUT, "gen"), exist_ok=True)
    path = os.path.join(OUT, "gen", "phold_b.cxx")
    body = model + PHOLD_B_MAIN
    if not os.path.exists(path) or open(path).read() != body:
        open(path, "w").write(body)
    return path


def phold_t_flags(cfg):
    root = "((uint64_t)((ray)/actor_n))" if cfg["rootdiv"] else "((uint64_t)(ray))"
    return [f"-DPHOLD_A={cfg['A']}", f"-DPHOLD_K={cfg['K']}", f"-DPHOLD_P={cfg['P']}",
            f"-DPHOLD_LAMBDA={cfg['LAMBDA']}", f"-DPHOLD_END={cfg['END']}",
            f"-DPHOLD_USER_SUBTIME={cfg['sub']}", f"-DPHOLD_ROOT_TIME(ray)={root}"]


def exe_name(kind, R, tag=None):
    if kind == "phold_t":
        return os.path.join(OUT, f"ref_phold_t_{tag}_R{R}")
    if kind == "phold_b":
        return os.path.join(OUT, f"ref_phold_b_R{R}")
    return os.path.join(OUT, f"ref_{kind}_R{R}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", type=int, nargs="*", default=[2, 8])
    ap.add_argument("--only", default="phold_test,phold_bench,phold_t,phold_b,heat,zero_delay,heat2d")
    ap.add_argument("--tags", default=",".join(PHOLD_T_CONFIGS))
    ap.add_argument("--baseline-ranks", type=int, nargs="*", default=[4, 16, 32, 64])
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    if not os.path.isdir(REF):
        print(f"oracle/build_ref.py: {REF} not present (GPU box?) - using prebuilt oracle/_ref/ as is")
        return 0
    os.makedirs(OUT, exist_ok=True)
    only = set(a.only.split(","))
    tags = [t for t in a.tags.split(",") if t]
    built = []
    for R in a.ranks:
        objs = build_runtime(R, a.force)
        jobs = []
        if "phold_test" in only:
            jobs.append((os.path.join(REF, "test/phold.cxx"), exe_name("phold_test", R), ()))
        if "phold_bench" in only:
            jobs.append((os.path.join(REF, "bench/phold.cxx"), exe_name("phold_bench", R),
                         (os.path.join(REF, "bench/util/report.cxx"),)))
        if "heat" in only:
            # example/heat (QSS actors), unchanged: QSS order 1 (its default) and order 0
            jobs.append((os.path.join(REF, "example/heat/heat.cxx"), exe_name("heat", R), ()))
            jobs.append((os.path.join(REF, "example/heat/heat.cxx"), exe_name("heat_q0", R), ("-DQSS_ORDER=0",)))
        if "zero_delay" in only:
            # tests/csrc/zero_delay.cxx: this repo's own test model (sends that keep the time stamp) on the reference runtime
            jobs.append((os.path.join(os.path.dirname(HERE), "tests", "csrc", "zero_delay.cxx"), exe_name("zero_delay", R), ()))
        if "heat2d" in only:
            # tests/csrc/heat2d.cxx: this repo's own 2-D heat stencil (BASELINE config 5's wording) on the reference runtime
            jobs.append((os.path.join(os.path.dirname(HERE), "tests", "csrc", "heat2d.cxx"), exe_name("heat2d", R), ()))
        if "phold_b" in only:
            jobs.append((gen_phold_b(), exe_name("phold_b", R), (os.path.join(REF, "bench/util/report.cxx"),)))
        if "phold_t" in only:
            gen = gen_phold_t()
            for tag in tags:
                if tag == "cfg3" and R != 8:
                    continue
                jobs.append((gen, exe_name("phold_t", R, tag), tuple(phold_t_flags(PHOLD_T_CONFIGS[tag]))))
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            built += list(ex.map(lambda j: link_app(R, objs, j[0], j[1], j[2], a.force), jobs))
    # the CPU-baseline binary (BASELINE config 2) for the host core counts a GPU box may have
    if "phold_t" in only:
        gen = gen_phold_t()
        for R in a.baseline_ranks:
            objs = build_runtime(R, a.force)
            for tag in BASELINE_TAGS:
                if tag in tags and not (R in a.ranks and tag == "cfg2"):
                    built.append(link_app(R, objs, gen, exe_name("phold_t", R, tag),
                                          tuple(phold_t_flags(PHOLD_T_CONFIGS[tag])), a.force))
    for b in built:
        print("built", os.path.relpath(b, os.path.dirname(HERE)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
