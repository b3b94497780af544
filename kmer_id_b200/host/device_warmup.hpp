// device_warmup.hpp - create the CUDA context on a helper thread while the host reads the probe
// file.  Context creation costs seconds on a large box; kid_db_build would otherwise pay it after
// the parse.  finish_device_warmup() joins (also at exit, so that no thread outlives main()).
#pragma once
#include <cstdlib>
#include <functional>
#include <thread>

#include "../../include/kmer_id.h"

namespace kidhost {

inline std::thread &warmup_thread()
{
    static std::thread t;
    return t;
}

inline void finish_device_warmup()
{
    if (warmup_thread().joinable()) warmup_thread().join();
}

// `then` (optional) runs on the helper thread once the context exists, e.g. page-locking buffers
inline void start_device_warmup(int device, std::function<void()> then = nullptr)
{
    if (warmup_thread().joinable()) return;
    atexit(finish_device_warmup);
    warmup_thread() = std::thread([device, then] {
        if (kid_device_init(device) == 0 && then) then(); // errors resurface in kid_db_build
    });
}

} // namespace kidhost
