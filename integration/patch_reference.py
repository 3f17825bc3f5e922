#!/usr/bin/env python
"""Generate the GPU-enabled copies of the four reference files the drop-in touches.

    python integration/patch_reference.py /root/reference/source/Lib  <scratch dir>

Reads the unmodified reference sources where they lie and writes patched copies to the scratch
directory (never into the repo): the patch is the one INTEGRATION.md documents --

  TLibCommon/TComRdCost.h      + three read-only accessors of the motion-cost state
  TLibEncoder/TEncSearch.cpp   bodies of xPatternSearch / xPatternSearchGT forward to hop_shim.h
  TLibEncoder/TEncCu.cpp       xCopyYuv2SSRef notifies the device mirror after the CU copy; xCompressCU opens a
                               speculation window in front of a CU's inter modes (first PU of every mode)
  TLibEncoder/TEncGOP.cpp      the mirror is (re)loaded when the SS reference is wired to the slice
  TLibEncoder/TEncTop.cpp      context created / destroyed with the encoder

Every edit is anchored on an exact line of the reference; the script fails loudly if an anchor moved.
"""
import os
import sys


def patch(text, anchor, insert, where="after", count=1):
    n = text.count(anchor)
    if n != count:
        raise SystemExit("anchor %r found %d times (expected %d)" % (anchor[:60], n, count))
    if where == "after":
        return text.replace(anchor, anchor + insert)
    if where == "replace":
        return text.replace(anchor, insert)
    return text.replace(anchor, insert + anchor)


def main():
    src, out = sys.argv[1], sys.argv[2]
    rd = lambda p: open(os.path.join(src, p), encoding="latin-1").read()

    def wr(p, t):
        os.makedirs(os.path.dirname(os.path.join(out, p)), exist_ok=True)
        open(os.path.join(out, p), "w", encoding="latin-1").write(t)

    # --- TComRdCost.h: accessors (inline, no layout change) ---------------------------------------
    t = rd("TLibCommon/TComRdCost.h")
    t = patch(t, "};// END CLASS DEFINITION TComRdCost",
              "public:\n"
              "  // libhopgpu: read-only view of the motion-cost state used by getCost()/getBits()\n"
              "  UInt          hopGetCost()      const { return m_uiCost; }\n"
              "  Int           hopGetCostScale() const { return m_iCostScale; }\n"
              "  const TComMv& hopGetPredictor() const { return m_mvPredictor; }\n", where="before")
    wr("TLibCommon/TComRdCost.h", t)

    # --- TEncSearch.cpp ---------------------------------------------------------------------------
    t = rd("TLibEncoder/TEncSearch.cpp")
    t = patch(t, '#include "TEncSearch.h"\n', '#include "hop_shim.h"   // libhopgpu\n')
    t = patch(t,
              "  UInt  uiSad;\n  UInt  uiSadBest         = MAX_UINT;\n  Int   iBestX = 0;\n  Int   iBestY = 0;\n  \n  Pel*  piRefSrch;\n",
              "  // libhopgpu: the SS full search runs on the GPU mirror of the SS reference\n"
              "  if ( isSSE && hopshim::owns( piRefY ) )\n"
              "  {\n"
              "    hopshim::xPatternSearch( pcPatternKey, piRefY, iRefStride, pcMvSrchRngLT, pcMvSrchRngRB, rcMv, ruiSAD,\n"
              "                             riOffsetX, riOffsetY, ssBestCand, isSSE, m_pcEncCfg->getUseFastEnc(), g_bitDepthY, m_pcRdCost );\n"
              "    return;\n"
              "  }\n", where="before")
    t = patch(t,
              "\t//gtFlag = true;\n\trcGT0->set(0,0);\n",
              "\t// libhopgpu: the HOP / geometric-transform candidate search runs on the GPU\n"
              "\tif ( hopshim::owns( piRefY ) )\n"
              "\t{\n"
              "\t\thopshim::xPatternSearchGT( pcCU, pcPatternKey, piRefY, iRefStride, pcMvInt, rcMvHalf, rcMvQter, rcGT0, rcGT1, rcGT2, rcGT3,\n"
              "\t\t                           gtFlag, ruiCost, bestSSCand, m_pcEncCfg->getUseHADME(), g_bitDepthY, m_pcRdCost );\n"
              "\t\treturn;\n"
              "\t}\n", where="before")
    # xMotionEstimation: the three SS stages in one GPU call; the original statements stay for every
    # other kind of reference
    t = patch(t, "  if ( !m_iFastSearch || bBi )\n  {\n#if IT_HOLOSS\n    if ( bIsSSE )\n    {\n      xSetSearchRange   (pcCU, cMvSrchRngLT, cMvSrchRngRB, iOffsetX, iOffsetY, bisFirstRow, bisFirstCol);\n    }\n",
              "  // libhopgpu: integer search, fractional refinement and HOP search of an SS reference in one GPU call\n"
              "  Bool bHopDone = false;\n"
              "  if ( bIsSSE && !bBi && !m_iFastSearch && hopshim::owns( piRefY ) && hopshim::fused() )\n"
              "  {\n"
              "    xSetSearchRange   (pcCU, cMvSrchRngLT, cMvSrchRngRB, iOffsetX, iOffsetY, bisFirstRow, bisFirstCol);\n"
              "    Bool bValid = hopshim::xMotionSearchSS( pcCU, pcPatternKey, piRefY, iRefStride, &cMvSrchRngLT, &cMvSrchRngRB, iOffsetX, iOffsetY,\n"
              "                                            m_pcEncCfg->getUseFastEnc(), m_pcEncCfg->getUseHADME(), g_bitDepthY, m_pcRdCost, bUseGT,\n"
              "                                            rcMv, ruiCost, cMvHalf, cMvQter, rcGT0, rcGT1, rcGT2, rcGT3, gtFlag, pcCU->getSSBestCand() );\n"
              "    if ( !bValid || pcCU->getSlice()->getRefPic( eRefPicList, iRefIdxPred )->getPicYuvRec()->getBufY()[0x00] == NOT_VALID )\n"
              "    {\n"
              "      bNotValCU = true;\n"
              "      return;\n"
              "    }\n"
              "    m_pcRdCost->getMotionCost( 1, 0 );\n"
              "    m_pcRdCost->setCostScale ( 0 );\n"
              "    bHopDone = true;\n"
              "  }\n"
              "  if ( !bHopDone )\n"
              "  {\n", where="before")
    t = patch(t, "  rcMv <<= 2;\n  rcMv += (cMvHalf <<= 1);\n", "  }  // libhopgpu: !bHopDone\n", where="before")
    wr("TLibEncoder/TEncSearch.cpp", t)

    # --- TEncCu.cpp -------------------------------------------------------------------------------
    t = rd("TLibEncoder/TEncCu.cpp")
    t = patch(t, '#include "TEncCu.h"\n', '#include "hop_shim.h"   // libhopgpu\n')
    # xCopyYuv2SSRef: the border re-extension and the device mirror update in one call (the two reference lines
    # stay reachable through hopshim::refCommit for planes that are not mirrored and with HOP_HOST_BORDER=full)
    t = patch(t,
              "    rpcPic->getPicYuvRec()->setBorderExtension(false);\n    rpcPic->getPicYuvRec()->extendPicBorder();\n",
              "    hopshim::refCommit( rpcPic->getPicYuvRec(), uiLPelX, uiTPelY, g_uiMaxCUWidth>>uiDepth, g_uiMaxCUHeight>>uiDepth );   // libhopgpu\n",
              where="replace")
    # speculation window: the reference's own xCheckRDCostInter is run once per partition mode
    # with the shim in "enqueue, do not wait" mode -- it derives the AMVP list and search window of the mode's
    # first PU exactly as the real pass will, hands the request to the GPU and returns through the reference's
    # own "no valid SS vector" exit; initEstData() then restores the CU as it does between any two modes.
    # ISS slices only: a PSS slice would run its temporal-reference search for real inside the window.
    t = patch(t,
              "      // do inter modes, SKIP and 2Nx2N\n      if( rpcBestCU->getSlice()->getSliceType() != I_SLICE )\n      {\n",
              "        // libhopgpu: enqueue the motion search of the first PU of every partition mode this CU is going to try\n"
              "        if ( hopshim::prefetchEnabled() && rpcBestCU->getSlice()->isIntraSS() )\n"
              "        {\n"
              "          static const PartSize aeHopModes[7] = { SIZE_2Nx2N, SIZE_Nx2N, SIZE_2NxN, SIZE_2NxnU, SIZE_2NxnD, SIZE_nLx2N, SIZE_nRx2N };\n"
              "          const Int iHopModes = ( hopshim::prefetchAmp() && pcPic->getSlice(0)->getSPS()->getAMPAcc(uiDepth) && rpcBestCU->getWidth(0) != 64 ) ? 7 : 3;\n"
              "          hopshim::prefetchBegin();\n"
              "          for ( Int iHopMode = 0; iHopMode < iHopModes; iHopMode++ )\n"
              "          {\n"
              "            xCheckRDCostInter( rpcBestCU, rpcTempCU, aeHopModes[iHopMode] );\n"
              "          }\n"
              "          hopshim::prefetchEnd();\n"
              "          rpcTempCU->initEstData( uiDepth, iQP, bIsLosslessMode );\n"
              "        }\n")
    wr("TLibEncoder/TEncCu.cpp", t)

    # --- TEncGOP.cpp ------------------------------------------------------------------------------
    t = rd("TLibEncoder/TEncGOP.cpp")
    t = patch(t, '#include "TEncGOP.h"\n', '#include "hop_shim.h"   // libhopgpu\n')
    t = patch(t,
              "    pcSlice->setRefPicList ( rcListPic, m_pcEncTop->getSSRefEncoder() );\n",
              "    // libhopgpu: the SS reference was reset and re-extended for an ISS / PSS slice: (re)load its device mirror\n"
              "    if ( pcSlice->isIntraSS() || pcSlice->isInterPSS() ) hopshim::refReset( m_pcEncTop->getSSRefEncoder()->getPicYuvRec() );\n"
              "    else hopshim::refInvalidate();\n")
    wr("TLibEncoder/TEncGOP.cpp", t)

    # --- TEncTop.cpp ------------------------------------------------------------------------------
    t = rd("TLibEncoder/TEncTop.cpp")
    t = patch(t, '#include "TEncTop.h"\n', '#include "hop_shim.h"   // libhopgpu\n')
    t = patch(t, "Void TEncTop::create ()\n{\n", "  hopshim::create();    // libhopgpu: one context per encoder instance\n")
    t = patch(t, "Void TEncTop::destroy ()\n{\n", "  hopshim::destroy();   // libhopgpu\n")
    wr("TLibEncoder/TEncTop.cpp", t)
    print("patched sources written to", out)


if __name__ == "__main__":
    main()
