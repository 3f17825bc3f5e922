// hop_batch_main.cpp -- batch front end of the GPU-backed encoder (BASELINE.json configs[3]: a queue of lenslet images per GPU).
//
// The reference's TAppEncoder encodes ONE sequence per process (source/App/TAppEncoder/encmain.cpp:53-100).  A CUDA
// context costs 0.3-2 s to create and the driver creates them one at a time machine-wide, so a
// process per image wastes most of a short encode when 32 encoders share an 8-GPU box.  This front end keeps the
// process -- and with it the CUDA context of libhopgpu -- alive and runs the reference's own TAppEncTop once per job,
// exactly as encmain.cpp does: create(), parseCfg(), encode(), destroy().  Nothing of the encoder proper is touched.
//
// Protocol (stdin -> stdout), one job per line, fields separated by TAB:
//     <working directory> TAB <arg1> TAB <arg2> ...        (the arguments TAppEncoder would get, without argv[0])
// answer, after the encoder's own output:   @@HOPBATCH done <rc> <seconds>
// An empty line or EOF ends the process.  Side files (TraceEnc.txt, cost.csv, psnr.txt, ...) land in the job's directory.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <string>
#include <vector>
#include <unistd.h>
#include "TAppEncTop.h"
#include "TLibEncoder/TEncAnalyze.h"
#include "TAppCommon/program_options_lite.h"
#ifndef HOP_BATCH_NO_GPU          /* -DHOP_BATCH_NO_GPU: the same front end over the UNPATCHED reference (CPU check of re-entrancy) */
#include "hop_shim.h"
#endif

namespace po = df::program_options_lite;

static double now_s()
{
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

static int run_job(std::vector<std::string>& f)
{
  if (chdir(f[0].c_str()) != 0) { fprintf(stderr, "hop_batch: cannot enter %s\n", f[0].c_str()); return 2; }
  std::vector<char*> argv;
  static char name[] = "TAppEncoderHopBatch";
  argv.push_back(name);
  for (size_t i = 1; i < f.size(); i++) argv.push_back(const_cast<char*>(f[i].c_str()));
  // process-wide state a fresh process would start with: the PSNR / bit-rate summary accumulators (TEncAnalyze.cpp:47-52;
  // they only feed the printed summary, TEncGOP.cpp:2138 asserts that they count this run's pictures)
  m_gcAnalyzeAll.clear(); m_gcAnalyzeI.clear(); m_gcAnalyzeP.clear(); m_gcAnalyzeB.clear(); m_gcAnalyzeAll_in.clear();
  TAppEncTop enc;                       // a fresh application object per job, as one process run would have
  enc.create();
  try {
    if (!enc.parseCfg((int)argv.size(), argv.data())) { enc.destroy(); return 1; }
  } catch (po::ParseFailure& e) {
    std::cerr << "Error parsing option \"" << e.arg << "\" with argument \"" << e.val << "\"." << std::endl;
    return 1;
  }
  const long before = clock();
  enc.encode();
  printf("\n Total Time: %12.3f sec.\n", (double)(clock() - before) / CLOCKS_PER_SEC);
  enc.destroy();
  return 0;
}

int main(int, char**)
{
  // the CUDA context is created here, once, before the first job arrives: a pool of workers pays the driver's serialised
  // context creation while its caller is still preparing input, not inside every image
  const double t0 = now_s();
#ifndef HOP_BATCH_NO_GPU
  hopshim::keepContext() = true;        // TEncTop::destroy of a job leaves the context (stream, pinned slots, scratch) alone
  hopshim::create();
  { hopshim::Stats& st = hopshim::stats(); for (int i = 0; i < 6; i++) { st.sec[i] = 0; st.calls[i] = 0; } }   // start-up is reported by "ready", not by the first job
#endif
  printf("@@HOPBATCH ready %.3f\n", now_s() - t0);
  fflush(stdout);
  std::string line;
  while (std::getline(std::cin, line)) {
    if (line.empty()) break;
    std::vector<std::string> f;
    size_t a = 0;
    while (true) {
      const size_t b = line.find('\t', a);
      f.push_back(line.substr(a, b == std::string::npos ? std::string::npos : b - a));
      if (b == std::string::npos) break;
      a = b + 1;
    }
    const double t1 = now_s();
    const int rc = run_job(f);
    fflush(stderr);
    // per-job share of the library calls (HOP_STATS=1): search calls incl. speculative enqueues, SS-mirror updates,
    // context / mirror creation -- then the counters start again for the next job
#ifndef HOP_BATCH_NO_GPU
    hopshim::Stats& st = hopshim::stats();
    printf("@@HOPBATCH stats %.4f %.4f %.4f\n", st.sec[0] + st.sec[1] + st.sec[4], st.sec[2] + st.sec[5], st.sec[3]);
    for (int i = 0; i < 6; i++) { st.sec[i] = 0; st.calls[i] = 0; }
#endif
    printf("@@HOPBATCH done %d %.3f\n", rc, now_s() - t1);
    fflush(stdout);
  }
#ifndef HOP_BATCH_NO_GPU
  hopshim::destroy(true);
#endif
  return 0;
}
