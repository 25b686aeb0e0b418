set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
cat > /tmp/t_par.cpp <<'EOC'
#include <thread>
#include <vector>
#include <chrono>
#include <cstdio>
#include <cstdint>
int main(){ for (int T : {1,2,4,8,16}) { auto t0=std::chrono::steady_clock::now(); std::vector<std::thread> th; std::vector<uint64_t> r(T);
 for (int t=0;t<T;++t) th.emplace_back([&,t]{ uint64_t x=t+1; for (uint64_t i=0;i<400000000ull/T;++i) x = x*6364136223846793005ull+1442695040888963407ull; r[t]=x;});
 for (auto&x:th) x.join(); auto t1=std::chrono::steady_clock::now(); printf("T=%d %.1f ms (%llu)\n",T,std::chrono::duration<double,std::milli>(t1-t0).count(),(unsigned long long)r[0]); } }
EOC
g++ -O2 -pthread /tmp/t_par.cpp -o /tmp/t_par && /tmp/t_par
timeout 900 python scripts/cli_e2e.py > gpurun_out/cli_e2e_c2.log 2>&1; echo "e2e exit=$?"; cat gpurun_out/cli_e2e_c2.log
SMAFA_HOST_THREADS=1 timeout 900 python scripts/cli_e2e.py > gpurun_out/cli_e2e_c2_1thread.log 2>&1; echo "e2e exit=$?"; grep -v "^$" gpurun_out/cli_e2e_c2_1thread.log | head -40
