"""Signatures of the network entry points (include/monopsr_b200_net.h); filled as they land."""


def declare(lib):
    return
