// stub for Jittor utils/log.h (Jittor is not installed); the reference headers only include it
