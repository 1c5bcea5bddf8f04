__global__ void ks(double* o, const double* a) { o[threadIdx.x] = sqrt(a[threadIdx.x]); }
