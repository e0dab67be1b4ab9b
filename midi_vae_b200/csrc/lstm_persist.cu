// lstm_persist.cu -- persistent LSTM recurrence kernels (placeholder until the tcgen05 version lands).
#include "common.cuh"
namespace mvae {
}
