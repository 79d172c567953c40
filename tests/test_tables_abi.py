"""CPU tests: generated constant tables vs the reference's, struct layouts, exported symbols."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import refharness as rh
from xeve_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built here")

PROBE = r'''
#include <stdio.h>
#include <stddef.h>
#include "include/xeve_b200.h"
#include "xeve_b200/csrc/xb200_tables.h"
int main(void){
  printf("seq %zu\nme %zu\nmc %zu\nrates %zu\ntq %zu\nres %zu\nblk %zu\n", sizeof(xb200_seq), sizeof(xb200_me_item),
         sizeof(xb200_mc_item), sizeof(xb200_rates), sizeof(xb200_tq_item), sizeof(xb200_residue_item), sizeof(xb200_blk_item));
  printf("sbac %zu\nbits %zu\ncu %zu\nmvpi %zu\n", sizeof(xb200_sbac), sizeof(xb200_bits_item), sizeof(xb200_cu_item), sizeof(xb200_mvp_item));
  printf("dfcu %zu\ndfpic %zu\ntrm %zu\n", sizeof(xb200_df_cu), sizeof(xb200_df_pic), sizeof(xb200_trm_item));
  printf("intra %zu\nnbr %zu\nintraoff %zu %zu %zu\n", sizeof(xb200_intra_item), sizeof(xb200_nbr_item), offsetof(xb200_intra_item, lambda),
         offsetof(xb200_intra_item, cost), offsetof(xb200_nbr_item, nb_off));
  { static const unsigned char m[6][6][5] = XB200_MPM_TABLE; FILE* g = fopen("mpm.bin","wb"); fwrite(m,1,sizeof(m),g); fclose(g); }
  printf("cuoff %zu %zu %zu %zu\n", offsetof(xb200_cu_item, lambda), offsetof(xb200_cu_item, mvp), offsetof(xb200_cu_item, cost), offsetof(xb200_bits_item, coef_off));
  printf("off %zu %zu %zu %zu\n", offsetof(xb200_me_item, lambda_mv), offsetof(xb200_tq_item, lambda),
         offsetof(xb200_residue_item, out_off), offsetof(xb200_residue_item, dist_rec));
  static int8_t tm[4096]; xb200_gen_tm64(tm);
  FILE* f = fopen("tm64.bin","wb"); fwrite(tm,1,4096,f); fclose(f);
  static uint16_t sc[4096];
  f = fopen("scan.bin","wb");
  for(int l=1;l<=6;l++){ xb200_gen_scan(sc,l,l); fwrite(sc,2,1<<(2*l),f);} fclose(f);
  f = fopen("mvbits.bin","wb"); for(int v=-2047; v<=2048; v++){ unsigned char b=(unsigned char)xb200_mvd_bits(v); fwrite(&b,1,1,f);} fclose(f);
  f = fopen("refi.bin","wb"); for(int n=0;n<17;n++) for(int r=0;r<16;r++){ unsigned char b = r<n||n==0 ? (unsigned char)xb200_refi_bits(n,r):0; fwrite(&b,1,1,f);} fclose(f);
  f = fopen("es.bin","wb"); for(int q=0;q<6;q++) for(int l=1;l<=7;l++){ long long e = xb200_err_scale(q,l,10); fwrite(&e,8,1,f);} fclose(f);
  f = fopen("dfst.bin","wb"); for(int i=0;i<4;i++) for(int q=0;q<52;q++){ unsigned char b=(unsigned char)xb200_df_strength(i,q); fwrite(&b,1,1,f);} fclose(f);
  return 0; }
'''


@pytest.fixture(scope="module")
def probe_dir():
    d = tempfile.mkdtemp()
    src = os.path.join(d, "probe.c")
    open(src, "w").write(PROBE)
    subprocess.check_call(["gcc", "-O1", "-I", ROOT, "-o", os.path.join(d, "probe"), src, "-lm"])
    out = subprocess.check_output([os.path.join(d, "probe")], cwd=d).decode()
    return d, out


def test_struct_layouts_match_numpy(probe_dir):
    _, out = probe_dir
    sizes = dict(re.findall(r"(\w+) (\d+)\n", out))
    assert int(sizes["seq"]) == api.SEQ.itemsize
    assert int(sizes["me"]) == api.ME_ITEM.itemsize
    assert int(sizes["mc"]) == api.MC_ITEM.itemsize
    assert int(sizes["rates"]) == api.RATES.itemsize
    assert int(sizes["tq"]) == api.TQ_ITEM.itemsize
    assert int(sizes["res"]) == api.RESIDUE_ITEM.itemsize
    assert int(sizes["blk"]) == api.BLK_ITEM.itemsize
    assert int(sizes["sbac"]) == api.SBAC.itemsize and int(sizes["bits"]) == api.BITS_ITEM.itemsize
    assert int(sizes["cu"]) == api.CU_ITEM.itemsize == rh.CU_REC.itemsize and int(sizes["mvpi"]) == api.MVP_ITEM.itemsize
    assert int(sizes["dfcu"]) == api.DF_CU.itemsize == rh.DF_CU.itemsize and int(sizes["dfpic"]) == api.DF_PIC.itemsize == rh.DF_PIC.itemsize
    assert int(sizes["intra"]) == api.INTRA_ITEM.itemsize == rh.INTRA_REC.itemsize and api.INTRA_ITEM.fields == rh.INTRA_REC.fields
    assert int(sizes["nbr"]) == api.NBR_ITEM.itemsize == rh.NBR_REC.itemsize
    assert int(sizes["trm"]) == api.TRM_ITEM.itemsize == 16 and api.TRM_ITEM.fields["off"][1] == 8
    ioffs = [int(v) for v in re.search(r"intraoff (\d+) (\d+) (\d+)", out).groups()]
    assert ioffs == [api.INTRA_ITEM.fields["lambda"][1], api.INTRA_ITEM.fields["cost"][1], api.NBR_ITEM.fields["nb_off"][1]]
    cuoffs = [int(v) for v in re.search(r"cuoff (\d+) (\d+) (\d+) (\d+)", out).groups()]
    assert cuoffs == [api.CU_ITEM.fields["lambda"][1], api.CU_ITEM.fields["mvp"][1], api.CU_ITEM.fields["cost"][1],
                      api.BITS_ITEM.fields["coef_off"][1]]
    assert api.CU_ITEM.fields == rh.CU_REC.fields and api.BITS_ITEM.fields == rh.BITS_REC.fields
    offs = [int(v) for v in re.search(r"\noff (\d+) (\d+) (\d+) (\d+)", out).groups()]
    assert offs == [api.ME_ITEM.fields["lambda_mv"][1], api.TQ_ITEM.fields["lambda"][1],
                    api.RESIDUE_ITEM.fields["out_off"][1], api.RESIDUE_ITEM.fields["dist_rec"][1]]
    # the harness records and the ABI records are the same bytes
    assert rh.ME_REC.itemsize == api.ME_ITEM.itemsize and rh.MC_REC.itemsize == api.MC_ITEM.itemsize
    assert rh.TQ_REC.itemsize == api.TQ_ITEM.itemsize and rh.RATES.itemsize == api.RATES.itemsize


@needs_ref
def test_generated_tables_match_reference(probe_dir):
    d, _ = probe_dir
    rd = lambda n, dt: np.fromfile(os.path.join(d, n), dt)
    assert np.array_equal(rd("tm64.bin", np.int8), rh.table(0, np.int8))
    tm64 = rh.table(0, np.int8).reshape(64, 64)
    for i, n in enumerate((32, 16, 8, 4, 2)):  # N-point matrices are sub-sampled rows of tm64
        assert np.array_equal(rh.table(1 + i, np.int8).reshape(n, n), tm64[:: 64 // n, :n])
    scan_ref = rh.table(8, np.uint16).reshape(6, 6, 4096)
    mine = rd("scan.bin", np.uint16)
    pos = 0
    for l in range(1, 7):
        n = 1 << (2 * l)
        assert np.array_equal(mine[pos:pos + n], scan_ref[l - 1, l - 1, :n]), l
        pos += n
    assert np.array_equal(rd("mvbits.bin", np.uint8), rh.table(6, np.uint8))
    assert np.array_equal(rd("refi.bin", np.uint8).reshape(17, 16), rh.table(7, np.uint8).reshape(17, 16))
    assert np.array_equal(rd("es.bin", np.int64).reshape(6, 7), rh.table(13, np.int64).reshape(6, 7))
    assert np.array_equal(rd("dfst.bin", np.uint8), rh.table(14, np.uint8))
    assert np.array_equal(rd("mpm.bin", np.uint8), rh.table(15, np.uint8))
    assert list(rh.table(9, np.int32)) == [40, 45, 51, 57, 64, 71]
    assert list(rh.table(10, np.int32)[:6]) == [26214, 23302, 20560, 18396, 16384, 14764]


ORACLE_PROBE = r"""
#include <stdio.h>
#include "oracle/xo_tables.h"
int main(void) {
  static int8_t tm[4096]; xo_gen_tm64(tm);
  FILE* f = fopen("o_tm64.bin","wb"); fwrite(tm,1,4096,f); fclose(f);
  static uint16_t sc[4096];
  f = fopen("o_scan.bin","wb"); for(int l=1;l<=6;l++){ xo_gen_scan(sc,l,l); fwrite(sc,2,1<<(2*l),f);} fclose(f);
  f = fopen("o_mvbits.bin","wb"); for(int v=-2047; v<=2048; v++){ unsigned char b=(unsigned char)xo_mvd_bits(v); fwrite(&b,1,1,f);} fclose(f);
  f = fopen("o_refi.bin","wb"); for(int n=0;n<17;n++) for(int r=0;r<16;r++){ unsigned char b = r<n||n==0 ? (unsigned char)xo_refi_bits(n,r):0; fwrite(&b,1,1,f);} fclose(f);
  f = fopen("o_es.bin","wb"); for(int q=0;q<6;q++) for(int l=1;l<=7;l++){ long long e = xo_err_scale(q,l,10); fwrite(&e,8,1,f);} fclose(f);
  f = fopen("o_dfst.bin","wb"); fwrite(xo_df_st,1,sizeof(xo_df_st),f); fclose(f);
  f = fopen("o_mpm.bin","wb"); fwrite(xo_mpm_tbl,1,sizeof(xo_mpm_tbl),f); fclose(f);
  f = fopen("o_taps.bin","wb"); fwrite(xo_mc_l_taps,1,sizeof(xo_mc_l_taps),f); fwrite(xo_mc_c_taps,1,sizeof(xo_mc_c_taps),f); fclose(f);
  f = fopen("o_q.bin","wb"); fwrite(xo_quant_scale,1,sizeof(xo_quant_scale),f); fwrite(xo_dequant_scale,1,sizeof(xo_dequant_scale),f); fclose(f);
  for(int v=-70000; v<=70000; v+=37) if(v > 2048 || v <= -2048) printf("esc %d %d\n", v, xo_mvd_bits(v));
  return 0; }
"""


@needs_ref
def test_oracle_tables_match_reference(probe_dir):
    """the oracle's OWN constants (oracle/xo_tables.h -- no code shared with the library's xb200_tables.h) against the reference's
    tables (src_base/xeve_tbl.c:40-48, 83-257, 286-517, 625-; src_base/xeve_mc.c:39-93; src_base/xeve_tq.c:37-39, 406-423), and the
    two generators against each other where the reference has no table (MVD escape lengths beyond +-2048)"""
    d = tempfile.mkdtemp()
    open(os.path.join(d, "oprobe.c"), "w").write(ORACLE_PROBE)
    subprocess.check_call(["gcc", "-O1", "-I", ROOT, "-o", os.path.join(d, "oprobe"), os.path.join(d, "oprobe.c"), "-lm"])
    out = subprocess.check_output([os.path.join(d, "oprobe")], cwd=d).decode()
    rd = lambda n, dt: np.fromfile(os.path.join(d, n), dt)
    assert np.array_equal(rd("o_tm64.bin", np.int8), rh.table(0, np.int8))
    scan_ref = rh.table(8, np.uint16).reshape(6, 6, 4096)
    mine, pos = rd("o_scan.bin", np.uint16), 0
    for l in range(1, 7):
        n = 1 << (2 * l)
        assert np.array_equal(mine[pos:pos + n], scan_ref[l - 1, l - 1, :n]), l
        pos += n
    assert np.array_equal(rd("o_mvbits.bin", np.uint8), rh.table(6, np.uint8))
    assert np.array_equal(rd("o_refi.bin", np.uint8).reshape(17, 16), rh.table(7, np.uint8).reshape(17, 16))
    assert np.array_equal(rd("o_es.bin", np.int64).reshape(6, 7), rh.table(13, np.int64).reshape(6, 7))
    assert np.array_equal(rd("o_dfst.bin", np.uint8), rh.table(14, np.uint8))
    assert np.array_equal(rd("o_mpm.bin", np.uint8), rh.table(15, np.uint8))
    taps = rd("o_taps.bin", np.int16)
    lref, cref = rh.table(11, np.int16).reshape(-1, 8), rh.table(12, np.int16).reshape(-1, 4)   # 1/16 and 1/32 sample phases
    assert np.array_equal(taps[:32].reshape(4, 8), lref[:: len(lref) // 4]) and np.array_equal(taps[32:].reshape(8, 4), cref[:: len(cref) // 8])
    q = rd("o_q.bin", np.int32)
    assert list(q[:6]) == list(rh.table(10, np.int32)[:6]) and list(q[6:]) == list(rh.table(9, np.int32))
    # escape branch: against the library's generator (compiled from ITS header)
    src = os.path.join(d, "esc.c")
    open(src, "w").write('#include <stdio.h>\n#include "xeve_b200/csrc/xb200_tables.h"\n'
                         'int main(void){ for(int v=-70000; v<=70000; v+=37) if(v > 2048 || v <= -2048) printf("esc %d %d\\n", v, xb200_mvd_bits(v)); return 0; }\n')
    subprocess.check_call(["gcc", "-O1", "-I", ROOT, "-o", os.path.join(d, "esc"), src, "-lm"])
    assert subprocess.check_output([os.path.join(d, "esc")]).decode() == out and out.count("esc") > 3000


def test_library_loads_and_exports_every_declared_symbol():
    """the C-ABI .so exists in-tree, dlopens without a GPU and exports what include/xeve_b200.h declares"""
    hdr = open(os.path.join(ROOT, "include", "xeve_b200.h")).read()
    declared = set(re.findall(r"XB200_API\s+[\w\s\*]+?\b(xb200_\w+)\s*\(", hdr))
    assert len(declared) >= 19
    L = api.load()
    for name in declared:
        assert hasattr(L, name), name
    assert set(api.EXPORTS) == declared
    assert b"sm_100a" in L.xb200_version()


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.Xb200Error) as e:
        api.Hotpath(api.make_seq(352, 288))
    assert e.value.code == api.ERR_UNSUPPORTED
