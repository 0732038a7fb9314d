// Probe (developer tool, not part of the product): which cross-process device-memory sharing mechanisms work on this
// box between two processes on two GPUs -- (1) legacy CUDA IPC (cudaIpcGetMemHandle / cudaIpcOpenMemHandle),
// (2) VMM allocations exported as POSIX file descriptors (cuMemCreate / cuMemExportToShareableHandle, the fd passed
// over a Unix socket). Build: nvcc -arch=sm_100a -o /tmp/ipc_probe tools/ipc_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <sys/socket.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

__global__ void k_fill(double* p, int n, double v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v + i;
  __threadfence_system();
}

static int send_fd(int sock, int fd) {
  char dummy = 'x';
  struct iovec io = {&dummy, 1};
  char ctl[CMSG_SPACE(sizeof(int))];
  std::memset(ctl, 0, sizeof(ctl));
  struct msghdr msg = {};
  msg.msg_iov = &io; msg.msg_iovlen = 1; msg.msg_control = ctl; msg.msg_controllen = sizeof(ctl);
  struct cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int));
  std::memcpy(CMSG_DATA(c), &fd, sizeof(int));
  return sendmsg(sock, &msg, 0) == 1 ? 0 : -1;
}
static int recv_fd(int sock) {
  char dummy;
  struct iovec io = {&dummy, 1};
  char ctl[CMSG_SPACE(sizeof(int))];
  struct msghdr msg = {};
  msg.msg_iov = &io; msg.msg_iovlen = 1; msg.msg_control = ctl; msg.msg_controllen = sizeof(ctl);
  if (recvmsg(sock, &msg, 0) != 1) return -1;
  struct cmsghdr* c = CMSG_FIRSTHDR(&msg);
  if (!c || c->cmsg_type != SCM_RIGHTS) return -1;
  int fd;
  std::memcpy(&fd, CMSG_DATA(c), sizeof(int));
  return fd;
}

int main() {
  int sv[2];
  if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) { perror("socketpair"); return 1; }
  const int n = 1 << 20;
  const pid_t pid = fork();
  if (pid == 0) {  // child: GPU 1 (or 0 if there is only one), the importer / writer
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    cudaSetDevice(ndev > 1 ? 1 : 0);
    cudaFree(0);
    // (1) legacy IPC
    cudaIpcMemHandle_t hd;
    if (read(sv[1], &hd, sizeof(hd)) != (ssize_t)sizeof(hd)) return 2;
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
    printf("child: cudaIpcOpenMemHandle: %s\n", cudaGetErrorString(e));
    int ok1 = 0;
    if (e == cudaSuccess) {
      k_fill<<<(n + 255) / 256, 256>>>((double*)p, n, 1000.0);
      e = cudaDeviceSynchronize();
      printf("child: kernel writing through the legacy mapping: %s\n", cudaGetErrorString(e));
      ok1 = e == cudaSuccess;
      cudaIpcCloseMemHandle(p);
    }
    cudaGetLastError();
    if (write(sv[1], &ok1, sizeof(int)) != (ssize_t)sizeof(int)) return 2;
    // (2) VMM + fd
    size_t size = 0;
    if (read(sv[1], &size, sizeof(size)) != (ssize_t)sizeof(size)) return 2;
    int ok2 = 0;
    if (size > 0) {
      const int fd = recv_fd(sv[1]);
      printf("child: received fd %d\n", fd);
      CUmemGenericAllocationHandle mh;
      CUresult r = cuMemImportFromShareableHandle(&mh, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
      printf("child: cuMemImportFromShareableHandle: %d\n", (int)r);
      if (r == CUDA_SUCCESS) {
        CUdeviceptr va = 0;
        r = cuMemAddressReserve(&va, size, 0, 0, 0);
        if (r == CUDA_SUCCESS) r = cuMemMap(va, size, 0, mh, 0);
        CUmemAccessDesc ad = {};
        int dev = 0;
        cudaGetDevice(&dev);
        ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ad.location.id = dev; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        if (r == CUDA_SUCCESS) r = cuMemSetAccess(va, size, &ad, 1);
        printf("child: reserve/map/set access: %d\n", (int)r);
        if (r == CUDA_SUCCESS) {
          k_fill<<<(n + 255) / 256, 256>>>((double*)va, n, 2000.0);
          e = cudaDeviceSynchronize();
          printf("child: kernel writing through the VMM mapping: %s\n", cudaGetErrorString(e));
          ok2 = e == cudaSuccess;
        }
      }
      close(fd);
    }
    if (write(sv[1], &ok2, sizeof(int)) != (ssize_t)sizeof(int)) return 2;
    return 0;
  }
  // parent: GPU 0, the exporter / reader
  cudaSetDevice(0);
  cudaFree(0);
  int can = -1, ndev = 0;
  cudaGetDeviceCount(&ndev);
  if (ndev > 1) cudaDeviceCanAccessPeer(&can, 1, 0);
  printf("parent: %d devices, device 1 can access device 0: %d\n", ndev, can);
  double* buf = nullptr;
  cudaMalloc((void**)&buf, sizeof(double) * n);
  cudaMemset(buf, 0, sizeof(double) * n);
  cudaDeviceSynchronize();
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, buf);
  printf("parent: cudaIpcGetMemHandle: %s\n", cudaGetErrorString(e));
  cudaGetLastError();
  if (write(sv[0], &hd, sizeof(hd)) != (ssize_t)sizeof(hd)) return 2;
  int ok1 = 0;
  if (read(sv[0], &ok1, sizeof(int)) != (ssize_t)sizeof(int)) return 2;
  double first[2] = {0, 0};
  cudaMemcpy(first, buf, sizeof(first), cudaMemcpyDeviceToHost);
  printf("parent: LEGACY IPC %s (buffer now starts with %.0f %.0f)\n", ok1 && first[0] == 1000.0 ? "WORKS" : "FAILS", first[0], first[1]);
  // VMM
  int dev = 0;
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = dev;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  CUresult r = cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
  size_t size = sizeof(double) * n;
  if (r == CUDA_SUCCESS && gran) size = (size + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle mh = 0;
  if (r == CUDA_SUCCESS) r = cuMemCreate(&mh, size, &prop, 0);
  printf("parent: cuMemCreate (granularity %zu): %d\n", gran, (int)r);
  int fd = -1;
  if (r == CUDA_SUCCESS) r = cuMemExportToShareableHandle(&fd, mh, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  printf("parent: cuMemExportToShareableHandle: %d (fd %d)\n", (int)r, fd);
  CUdeviceptr va = 0;
  if (r == CUDA_SUCCESS) {
    r = cuMemAddressReserve(&va, size, 0, 0, 0);
    if (r == CUDA_SUCCESS) r = cuMemMap(va, size, 0, mh, 0);
    CUmemAccessDesc ad = {};
    ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ad.location.id = dev; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (r == CUDA_SUCCESS) r = cuMemSetAccess(va, size, &ad, 1);
    if (r == CUDA_SUCCESS) cudaMemset((void*)va, 0, size);
    cudaDeviceSynchronize();
  }
  size_t send_size = r == CUDA_SUCCESS ? size : 0;
  if (write(sv[0], &send_size, sizeof(send_size)) != (ssize_t)sizeof(send_size)) return 2;
  int ok2 = 0;
  if (send_size) {
    send_fd(sv[0], fd);
  }
  if (read(sv[0], &ok2, sizeof(int)) != (ssize_t)sizeof(int)) return 2;
  if (send_size) {
    cudaMemcpy(first, (void*)va, sizeof(first), cudaMemcpyDeviceToHost);
    printf("parent: VMM + fd %s (buffer now starts with %.0f %.0f)\n", ok2 && first[0] == 2000.0 ? "WORKS" : "FAILS", first[0], first[1]);
  } else {
    printf("parent: VMM + fd FAILS (export)\n");
  }
  int st = 0;
  waitpid(pid, &st, 0);
  return 0;
}
