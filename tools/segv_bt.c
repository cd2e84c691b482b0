// LD_PRELOAD helper (development aid): prints a native backtrace on SIGSEGV.
#define _GNU_SOURCE
#include <execinfo.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
static void handler(int sig) {
    void *frames[64];
    int n = backtrace(frames, 64);
    fprintf(stderr, "=== signal %d, native backtrace ===\n", sig);
    backtrace_symbols_fd(frames, n, 2);
    _exit(139);
}
__attribute__((constructor)) static void install(void) {
    struct sigaction sa;
    sa.sa_handler = handler;
    sigemptyset(&sa.sa_mask);
    sa.sa_flags = SA_RESETHAND;
    sigaction(SIGSEGV, &sa, NULL);
}
