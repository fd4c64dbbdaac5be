// Host-side launcher for the wavefront kernels (see wavefront.cuh).
#pragma once

#include "sim.h"
#include "wavefront.cuh"

template <class Op, int SX, int SY>
int launchWavefront(Sim* s, const Op& op, int ncb, int nstrips, const int* gate, int gateRunIfNonZero, int* changed) {
    wf::Domain dom{ncb, nstrips, s->fr.pitch};
    wf::Control ctl{s->wfTicket, s->wfFinished, s->hand, gate, gateRunIfNonZero, changed, s->dbgState};
    if ((size_t)Op::W * nstrips * ncb * 32 > s->handWords) {
        fsim_set_error("wavefront hand-off buffer too small");
        return FSIM_E_STATE;
    }
    if (s->opt.debugSimpleWavefront) {
        wf::wavefrontSimpleKernel<Op, SX, SY><<<1, 1024, 0, s->stream>>>(op, dom, ctl);
    } else {
        static bool attrSet[16] = {};
        size_t bytes = wf::Layout<Op>::BYTES;
        if (!attrSet[s->device & 15]) {
            CUDA_TRY(cudaFuncSetAttribute(wf::wavefrontKernel<Op, SX, SY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            attrSet[s->device & 15] = true;
        }
        wf::wavefrontKernel<Op, SX, SY><<<nstrips, 96, bytes, s->stream>>>(op, dom, ctl);
    }
    LAUNCH_COUNT(s);
    CUDA_TRY(cudaGetLastError());
    return FSIM_OK;
}
