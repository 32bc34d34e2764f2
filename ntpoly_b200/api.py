"""Python host side: a ctypes mirror of NTPoly's SWIG module (`NTPolySwig`) over the
C ABI of libntpoly_b200.so.

Class and method names follow the reference C++/SWIG classes
(reference Source/CPlusPlus/PSMatrix.h, SolverParameters.h, DensityMatrixSolvers.h,
SignSolvers.h, InverseSolvers.h, SquareRootSolvers.h, ExponentialSolvers.h,
EigenBounds.h, LoadBalancer.h, Permutation.h, PMatrixMemoryPool.h, TripletList.h,
ProcessGrid.h) so that the reference's own test scripts read the same:

    import ntpoly_b200.api as nt
    nt.ConstructGlobalProcessGrid(1, 1, 1)
    A = nt.Matrix_ps(n); A.FillFromTripletList(tl)
    C = nt.Matrix_ps(n); pool = nt.PMatrixMemoryPool(A)
    C.Gemm(A, A, pool, threshold=1e-8)

There is NO CPU fallback: every numerical call goes through the CUDA library and the
process aborts if no GPU is visible.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_bool, c_char_p, c_double, c_int, c_long, c_longlong, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NTB_LIB") or os.path.join(_HERE, "lib", "libntpoly_b200.so")   # (NTB_LIB: A/B runs of a kernel variant)
SIZE_wrp = 12
_lib = None

# every symbol include/ntpoly_b200.h declares (used by the "library exports" test)
EXPORTED_SYMBOLS = """
ConstructGlobalProcessGrid_wrp ConstructGlobalProcessGrid_onlyslice_wrp ConstructGlobalProcessGrid_default_wrp
CopyProcessGrid_wrp GetGlobalMySlice_wrp GetGlobalMyColumn_wrp GetGlobalMyRow_wrp GetGlobalIsRoot_wrp
GetGlobalNumSlices_wrp GetGlobalNumColumns_wrp GetGlobalNumRows_wrp WriteGlobalProcessGridInfo_wrp
DestructGlobalProcessGrid_wrp ConstructProcessGrid_wrp ConstructProcessGrid_onlyslice_wrp
ConstructProcessGrid_default_wrp GetMySlice_wrp GetMyColumn_wrp GetMyRow_wrp GetNumSlices_wrp GetNumColumns_wrp
GetNumRows_wrp WriteProcessGridInfo_wrp DestructProcessGrid_wrp
ConstructTripletList_r_wrp ResizeTripletList_r_wrp AppendToTripletList_r_wrp SetTripletAt_r_wrp GetTripletAt_r_wrp
DestructTripletList_r_wrp GetTripletListSize_r_wrp ConstructTripletList_c_wrp ResizeTripletList_c_wrp
AppendToTripletList_c_wrp SetTripletAt_c_wrp GetTripletAt_c_wrp DestructTripletList_c_wrp GetTripletListSize_c_wrp
ConstructEmptyMatrix_ps_wrp ConstructEmptyMatrixPG_ps_wrp CopyMatrix_ps_wrp DestructMatrix_ps_wrp
ConstructMatrixFromMatrixMarket_ps_wrp ConstructMatrixFromMatrixMarketPG_ps_wrp WriteMatrixToMatrixMarket_ps_wrp
FillMatrixFromTripletList_psr_wrp FillMatrixFromTripletList_psc_wrp FillMatrixPermutation_ps_wrp
FillMatrixIdentity_ps_wrp GetMatrixActualDimension_ps_wrp GetMatrixLogicalDimension_ps_wrp GetMatrixSize_ps_wrp
GetMatrixTripletList_psr_wrp GetMatrixTripletList_psc_wrp TransposeMatrix_ps_wrp ConjugateMatrix_ps_wrp
GetMatrixProcessGrid_ps_wrp IsIdentity_ps_wrp FillMatrixDense_ps_wrp GetMatrixBlock_psr_wrp GetMatrixBlock_psc_wrp
GetMatrixSlice_wrp ResizeMatrix_ps_wrp SnapMatrixToSparsityPattern_wrp
MatrixMultiply_ps_wrp IncrementMatrix_ps_wrp ScaleMatrix_ps_wrp MatrixTrace_ps_wrp MatrixNorm_ps_wrp
DotMatrix_psr_wrp DotMatrix_psc_wrp MatrixPairwiseMultiply_ps_wrp MeasureAsymmetry_ps_wrp SymmetrizeMatrix_ps_wrp
ConstructMatrixMemoryPool_p_wrp DestructMatrixMemoryPool_p_wrp
ConstructSolverParameters_wrp SetParametersConvergeDiff_wrp SetParametersMaxIterations_wrp
SetParametersBeVerbose_wrp SetParametersThreshold_wrp SetParametersLoadBalance_wrp SetParametersStepThreshold_wrp
SetParametersMonitorConvergence_wrp DestructSolverParameters_wrp ConstructDefaultPermutation_wrp
ConstructReversePermutation_wrp ConstructRandomPermutation_wrp DestructPermutation_wrp PermuteMatrix_wrp
UndoPermuteMatrix_wrp
TRS2_wrp TRS4_wrp PM_wrp HPCP_wrp EnergyDensityMatrix_wrp McWeenyStep_wrp McWeenyStepS_wrp SignFunction_wrp
PolarDecomposition_wrp Invert_wrp SquareRoot_wrp InverseSquareRoot_wrp ComputeExponential_wrp GershgorinBounds_wrp
PowerBounds_wrp
SortTripletList_r_wrp SortTripletList_c_wrp MatrixDiagonalScale_psr_wrp MatrixDiagonalScale_psc_wrp
ScaleAndFold_wrp PseudoInverse_wrp ActivateLogger_wrp ActivateLoggerFile_wrp DeactivateLogger_wrp
ConstructMatrixFromFile_lsr_wrp ConstructMatrixFromTripletList_lsr_wrp ConstructZeroMatrix_lsr_wrp
DestructMatrix_lsr_wrp CopyMatrix_lsr_wrp GetMatrixRows_lsr_wrp GetMatrixColumns_lsr_wrp ExtractMatrixRow_lsr_wrp
ExtractMatrixColumn_lsr_wrp ScaleMatrix_lsr_wrp IncrementMatrix_lsr_wrp DotMatrix_lsr_wrp
PairwiseMultiplyMatrix_lsr_wrp MatrixMultiply_lsr_wrp TransposeMatrix_lsr_wrp PrintMatrix_lsr_wrp
PrintMatrixF_lsr_wrp MatrixToTripletList_lsr_wrp MatrixDiagonalScale_lsr_wrp
ConstructMatrixFromFile_lsc_wrp ConstructMatrixFromTripletList_lsc_wrp ConstructZeroMatrix_lsc_wrp
DestructMatrix_lsc_wrp CopyMatrix_lsc_wrp GetMatrixRows_lsc_wrp GetMatrixColumns_lsc_wrp ExtractMatrixRow_lsc_wrp
ExtractMatrixColumn_lsc_wrp ScaleMatrix_lsc_wrp IncrementMatrix_lsc_wrp DotMatrix_lsc_wrp
PairwiseMultiplyMatrix_lsc_wrp MatrixMultiply_lsc_wrp TransposeMatrix_lsc_wrp ConjugateMatrix_lsc_wrp
PrintMatrix_lsc_wrp PrintMatrixF_lsc_wrp MatrixToTripletList_lsc_wrp MatrixDiagonalScale_lsc_wrp
ConstructMatrixMemoryPool_lr_wrp DestructMatrixMemoryPool_lr_wrp ConstructMatrixMemoryPool_lc_wrp
DestructMatrixMemoryPool_lc_wrp
ntb_nccl_unique_id ntb_world_init ntb_world_rank ntb_world_size ntb_set_stream ntb_synchronize
ntb_TripletList_r_set ntb_TripletList_r_get ntb_TripletList_c_set ntb_TripletList_c_get
ntb_FillMatrixFromArrays_ps ntb_GetMatrixLocalSize_ps ntb_GetMatrixArrays_ps ntb_GetMatrixArraysAsync_ps ntb_EgressWait ntb_StageArrays ntb_FillMatrixFromStaged_ps ntb_sorted_ingests ntb_ConstructEmptyMatrixComplex_ps
ntb_MatrixIsComplex_ps ntb_FilterMatrix_ps ntb_ScaleMatrixComplex_ps ntb_InverseSquareRootOrder_wrp
ntb_SquareRootOrder_wrp ntb_ConstructRandomPermutationSeeded ntb_SetPermutation ntb_get_counters
ntb_set_fused_shift ntb_set_fused_norm ntb_fused_norms ntb_get_halo_counters ntb_get_peer_counters ntb_measure_dmma_peak_tflops ntb_get_sync_count ntb_profile_read_phases ntb_set_halo_path ntb_set_permute_gemm ntb_TileCombine_ps ntb_TileScalars_ps ntb_set_fused_steps ntb_tile_combines ntb_hash_columns ntb_complex_tile_products ntb_tile_builds ntb_MatrixMultiplyShift_ps ntb_SignIteration ntb_SignStep ntb_set_flop_counting ntb_get_deferred_counters ntb_grid_layout ntb_default_grid ntb_reset_counters ntb_get_tile_counters ntb_set_tile_path ntb_algorithmic_bytes ntb_profile_enable ntb_profile_read ntb_last_solve ntb_MatrixAlgorithmicBytes_ps ntb_version
""".split()


def lib():
    """Load libntpoly_b200.so (built in-tree by `python -m ntpoly_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m ntpoly_b200.build` "
                "(there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        L.MatrixNorm_ps_wrp.restype = c_double
        L.ntb_tile_builds.restype = c_double
        L.ntb_sorted_ingests.restype = c_double
        L.ntb_SignIteration.restype = c_double
        L.ntb_SignStep.restype = c_double
        L.MeasureAsymmetry_ps_wrp.restype = c_double
        L.GetGlobalIsRoot_wrp.restype = c_bool
        L.ntb_GetMatrixLocalSize_ps.restype = c_longlong
        L.ntb_GetMatrixArraysAsync_ps.restype = c_longlong
        L.ntb_MatrixAlgorithmicBytes_ps.restype = c_longlong
        L.ntb_version.restype = c_char_p
        L.ntb_algorithmic_bytes.restype = c_double
        L.ntb_measure_dmma_peak_tflops.restype = c_double
        L.ntb_get_sync_count.restype = c_double
        L.ntb_tile_combines.restype = c_double
        L.ntb_hash_columns.restype = c_double
        L.ntb_fused_norms.restype = c_double
        L.ntb_complex_tile_products.restype = c_double
        L.ntb_set_stream.argtypes = [c_void_p]
        L.ntb_world_init.argtypes = [c_int, c_int, c_void_p]
        L.ntb_nccl_unique_id.argtypes = [c_void_p]
        _lib = L
    return _lib


def _handle():
    return (c_int * SIZE_wrp)()


def _d(x):
    return byref(c_double(float(x)))


def _i(x):
    return byref(c_int(int(x)))


def _ip(a):
    return a.ctypes.data_as(POINTER(c_int))


def _dp(a):
    return a.ctypes.data_as(POINTER(c_double))


# ---------------------------------------------------------------------------
# bootstrap / process grid (reference Source/CPlusPlus/ProcessGrid.h free functions)
# ---------------------------------------------------------------------------
def init_world_from_torch():
    """One process per GPU under torchrun: broadcast the ncclUniqueId with
    torch.distributed (any backend) and hand it to the library."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    size = int(os.environ.get("WORLD_SIZE", "1"))
    if size == 1:
        lib().ntb_world_init(0, 1, None)
        return 0, 1
    if not dist.is_initialized():
        dist.init_process_group("gloo")
    idbuf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        raw = (ctypes.c_ubyte * 128)()
        lib().ntb_nccl_unique_id(raw)
        idbuf = torch.tensor(list(raw), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        dev = idbuf.cuda()
        dist.broadcast(dev, 0)
        idbuf = dev.cpu()
    else:
        dist.broadcast(idbuf, 0)
    raw = (ctypes.c_ubyte * 128)(*idbuf.tolist())
    lib().ntb_world_init(rank, size, raw)
    return rank, size


def ConstructGlobalProcessGrid(rows=None, cols=None, slices=None, be_verbose=False):
    comm = c_int(0)
    if rows is None:
        lib().ConstructGlobalProcessGrid_default_wrp(byref(comm))
    elif cols is None:
        lib().ConstructGlobalProcessGrid_onlyslice_wrp(byref(comm), _i(rows))
    else:
        lib().ConstructGlobalProcessGrid_wrp(byref(comm), _i(rows), _i(cols), _i(slices if slices else 1))
    if be_verbose:
        lib().WriteGlobalProcessGridInfo_wrp()


def DestructGlobalProcessGrid():
    lib().DestructGlobalProcessGrid_wrp()


def GetGlobalMySlice():
    return lib().GetGlobalMySlice_wrp()


def GetGlobalMyColumn():
    return lib().GetGlobalMyColumn_wrp()


def GetGlobalMyRow():
    return lib().GetGlobalMyRow_wrp()


def GetGlobalIsRoot():
    return bool(lib().GetGlobalIsRoot_wrp())


def GetGlobalNumSlices():
    return lib().GetGlobalNumSlices_wrp()


def GetGlobalNumColumns():
    return lib().GetGlobalNumColumns_wrp()


def GetGlobalNumRows():
    return lib().GetGlobalNumRows_wrp()


def WriteGridInfo():
    lib().WriteGlobalProcessGridInfo_wrp()


def ActivateLogger(file_name=None, start_document=False):
    """reference Source/CPlusPlus/Logging.h; the YAML logger itself is outside the path (accepted, no output)"""
    if file_name is None or isinstance(file_name, bool):
        lib().ActivateLogger_wrp(byref(c_bool(bool(file_name) if isinstance(file_name, bool) else start_document)))
    else:
        b = file_name.encode()
        lib().ActivateLoggerFile_wrp(byref(c_bool(start_document)), c_char_p(b), _i(len(b)))


def DeactivateLogger():
    lib().DeactivateLogger_wrp()


class ProcessGrid:
    def __init__(self, rows=None, cols=None, slices=None):
        self.ih = _handle()
        comm = c_int(0)
        if rows is None:
            lib().ConstructProcessGrid_default_wrp(self.ih, byref(comm))
        elif cols is None:
            lib().ConstructProcessGrid_onlyslice_wrp(self.ih, byref(comm), _i(rows))
        else:
            lib().ConstructProcessGrid_wrp(self.ih, byref(comm), _i(rows), _i(cols), _i(slices if slices else 1))

    def GetMySlice(self):
        return lib().GetMySlice_wrp(self.ih)

    def GetMyColumn(self):
        return lib().GetMyColumn_wrp(self.ih)

    def GetMyRow(self):
        return lib().GetMyRow_wrp(self.ih)

    def GetNumSlices(self):
        return lib().GetNumSlices_wrp(self.ih)

    def GetNumColumns(self):
        return lib().GetNumColumns_wrp(self.ih)

    def GetNumRows(self):
        return lib().GetNumRows_wrp(self.ih)

    def WriteInfo(self):
        lib().WriteProcessGridInfo_wrp(self.ih)


# ---------------------------------------------------------------------------
# triplets
# ---------------------------------------------------------------------------
class Triplet_r:
    def __init__(self, index_column=0, index_row=0, point_value=0.0):
        self.index_column, self.index_row, self.point_value = index_column, index_row, point_value


class Triplet_c:
    def __init__(self, index_column=0, index_row=0, point_value=0j):
        self.index_column, self.index_row, self.point_value = index_column, index_row, point_value


class TripletList_r:
    is_complex = False

    def __init__(self, size=0):
        self.ih = _handle()
        lib().ConstructTripletList_r_wrp(self.ih, _i(size))

    def __del__(self):
        try:
            lib().DestructTripletList_r_wrp(self.ih)
        except Exception:
            pass

    def Resize(self, size):
        lib().ResizeTripletList_r_wrp(self.ih, _i(size))

    def Append(self, t):
        lib().AppendToTripletList_r_wrp(self.ih, _i(t.index_column), _i(t.index_row), _d(t.point_value))

    def SetTripletAt(self, index, t):
        lib().SetTripletAt_r_wrp(self.ih, _i(index + 1), _i(t.index_column), _i(t.index_row), _d(t.point_value))

    def GetTripletAt(self, index):
        c, r, v = c_int(), c_int(), c_double()
        lib().GetTripletAt_r_wrp(self.ih, _i(index + 1), byref(c), byref(r), byref(v))
        return Triplet_r(c.value, r.value, v.value)

    def GetSize(self):
        return lib().GetTripletListSize_r_wrp(self.ih)

    @staticmethod
    def SortTripletList(input_list, matrix_columns, sorted_list):
        """sorted copy of input_list (by column, then row) into sorted_list (reference TripletList.h:42-43)"""
        lib().DestructTripletList_r_wrp(sorted_list.ih)      # the C entry point allocates a fresh list for the result
        lib().SortTripletList_r_wrp(input_list.ih, _i(matrix_columns), sorted_list.ih)

    def Sort(self, matrix_columns=0):
        out = TripletList_r()
        TripletList_r.SortTripletList(self, matrix_columns, out)
        return out

    # bulk (extension)
    def set_arrays(self, rows, cols, vals):
        rows = np.ascontiguousarray(rows, np.int32)
        cols = np.ascontiguousarray(cols, np.int32)
        vals = np.ascontiguousarray(vals, np.float64)
        lib().ntb_TripletList_r_set(self.ih, c_longlong(len(rows)), _ip(rows), _ip(cols), _dp(vals))

    def get_arrays(self):
        n = self.GetSize()
        rows, cols, vals = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float64)
        lib().ntb_TripletList_r_get(self.ih, _ip(rows), _ip(cols), _dp(vals))
        return rows, cols, vals


class TripletList_c:
    is_complex = True

    def __init__(self, size=0):
        self.ih = _handle()
        lib().ConstructTripletList_c_wrp(self.ih, _i(size))

    def __del__(self):
        try:
            lib().DestructTripletList_c_wrp(self.ih)
        except Exception:
            pass

    def Append(self, t):
        v = complex(t.point_value)
        lib().AppendToTripletList_c_wrp(self.ih, _i(t.index_column), _i(t.index_row), _d(v.real), _d(v.imag))

    def GetTripletAt(self, index):
        c, r, re, im = c_int(), c_int(), c_double(), c_double()
        lib().GetTripletAt_c_wrp(self.ih, _i(index + 1), byref(c), byref(r), byref(re), byref(im))
        return Triplet_c(c.value, r.value, complex(re.value, im.value))

    def GetSize(self):
        return lib().GetTripletListSize_c_wrp(self.ih)

    def Resize(self, size):
        lib().ResizeTripletList_c_wrp(self.ih, _i(size))

    def SetTripletAt(self, index, t):
        v = complex(t.point_value)
        lib().SetTripletAt_c_wrp(self.ih, _i(index + 1), _i(t.index_column), _i(t.index_row), _d(v.real), _d(v.imag))

    @staticmethod
    def SortTripletList(input_list, matrix_columns, sorted_list):
        lib().DestructTripletList_c_wrp(sorted_list.ih)
        lib().SortTripletList_c_wrp(input_list.ih, _i(matrix_columns), sorted_list.ih)

    def Sort(self, matrix_columns=0):
        out = TripletList_c()
        TripletList_c.SortTripletList(self, matrix_columns, out)
        return out

    def set_arrays(self, rows, cols, vals):
        rows = np.ascontiguousarray(rows, np.int32)
        cols = np.ascontiguousarray(cols, np.int32)
        vals = np.ascontiguousarray(vals, np.complex128)
        lib().ntb_TripletList_c_set(self.ih, c_longlong(len(rows)), _ip(rows), _ip(cols),
                                    vals.view(np.float64).ctypes.data_as(POINTER(c_double)))

    def get_arrays(self):
        n = self.GetSize()
        rows, cols, vals = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.complex128)
        lib().ntb_TripletList_c_get(self.ih, _ip(rows), _ip(cols),
                                    vals.view(np.float64).ctypes.data_as(POINTER(c_double)))
        return rows, cols, vals


# ---------------------------------------------------------------------------
class Permutation:
    def __init__(self, matrix_dimension):
        self.n = int(matrix_dimension)
        self.ih = _handle()
        lib().ConstructDefaultPermutation_wrp(self.ih, _i(self.n))

    def _replace(self, fn, *args):
        lib().DestructPermutation_wrp(self.ih)
        self.ih = _handle()
        fn(self.ih, _i(self.n), *args)

    def SetDefaultPermutation(self):
        self._replace(lib().ConstructDefaultPermutation_wrp)

    def SetReversePermutation(self):
        self._replace(lib().ConstructReversePermutation_wrp)

    def SetRandomPermutation(self, seed=None):
        if seed is None:
            self._replace(lib().ConstructRandomPermutation_wrp)
        else:
            self._replace(lib().ntb_ConstructRandomPermutationSeeded, byref(c_longlong(int(seed))))

    def SetLookup(self, index_lookup_1based):
        a = np.ascontiguousarray(index_lookup_1based, np.int32)
        self._replace(lib().ntb_SetPermutation, _ip(a))


class SolverParameters:
    def __init__(self):
        self.ih = _handle()
        lib().ConstructSolverParameters_wrp(self.ih)

    def SetConvergeDiff(self, v):
        lib().SetParametersConvergeDiff_wrp(self.ih, _d(v))

    def SetMaxIterations(self, v):
        lib().SetParametersMaxIterations_wrp(self.ih, _i(v))

    def SetVerbosity(self, v):
        lib().SetParametersBeVerbose_wrp(self.ih, byref(c_bool(bool(v))))

    def SetThreshold(self, v):
        lib().SetParametersThreshold_wrp(self.ih, _d(v))

    def SetLoadBalance(self, perm: Permutation):
        lib().SetParametersLoadBalance_wrp(self.ih, perm.ih)

    def SetStepThreshold(self, v):
        lib().SetParametersStepThreshold_wrp(self.ih, _d(v))

    def SetMonitorConvergence(self, v):
        lib().SetParametersMonitorConvergence_wrp(self.ih, byref(c_bool(bool(v))))


class PMatrixMemoryPool:
    def __init__(self, matrix):
        self.ih = _handle()
        lib().ConstructMatrixMemoryPool_p_wrp(self.ih, matrix.ih)

    def __del__(self):
        try:
            lib().DestructMatrixMemoryPool_p_wrp(self.ih)
        except Exception:
            pass


# ---------------------------------------------------------------------------
# local matrices (reference Source/CPlusPlus/SMatrix.h, MatrixMemoryPool.h)
# ---------------------------------------------------------------------------
class _MatrixMemoryPool:
    _sfx = "lr"

    def __init__(self, columns, rows):
        self.ih = _handle()
        getattr(lib(), f"ConstructMatrixMemoryPool_{self._sfx}_wrp")(self.ih, _i(columns), _i(rows))

    def __del__(self):
        try:
            getattr(lib(), f"DestructMatrixMemoryPool_{self._sfx}_wrp")(self.ih)
        except Exception:
            pass


class MatrixMemoryPool_r(_MatrixMemoryPool):
    _sfx = "lr"


class MatrixMemoryPool_c(_MatrixMemoryPool):
    _sfx = "lc"


class _Matrix_ls:
    """Local sparse matrix living in GPU memory as one CSC block; same constructor overloads and method names as the
    reference class: (columns, rows) | (file_name) | (triplet_list, rows, columns) | (other matrix)."""
    _sfx = "lsr"
    is_complex = False

    def _fn(self, name):
        return getattr(lib(), f"{name}_{self._sfx}_wrp")

    def __init__(self, *args):
        self.ih = _handle()
        if len(args) == 1 and isinstance(args[0], _Matrix_ls):
            other = args[0]
            self._fn("ConstructZeroMatrix")(self.ih, _i(other.GetRows()), _i(other.GetColumns()))
            self._fn("CopyMatrix")(other.ih, self.ih)
        elif len(args) == 1 and isinstance(args[0], str):
            b = args[0].encode()
            self._fn("ConstructMatrixFromFile")(self.ih, c_char_p(b), _i(len(b)))
        elif len(args) == 3:
            tl, rows, columns = args
            self._fn("ConstructMatrixFromTripletList")(self.ih, tl.ih, _i(rows), _i(columns))
        else:
            columns, rows = args
            self._fn("ConstructZeroMatrix")(self.ih, _i(rows), _i(columns))

    def __del__(self):
        try:
            self._fn("DestructMatrix")(self.ih)
        except Exception:
            pass

    def GetRows(self):
        v = c_int()
        self._fn("GetMatrixRows")(self.ih, byref(v))
        return v.value

    def GetColumns(self):
        v = c_int()
        self._fn("GetMatrixColumns")(self.ih, byref(v))
        return v.value

    def ExtractRow(self, row_number, row_out):
        """row_number is 0-based like in the reference C++ class; row_out receives a 1 x columns matrix"""
        self._fn("DestructMatrix")(row_out.ih)       # the C entry point allocates a fresh object for the output
        self._fn("ExtractMatrixRow")(self.ih, byref(c_int(row_number + 1)), row_out.ih)

    def ExtractColumn(self, column_number, column_out):
        self._fn("DestructMatrix")(column_out.ih)
        self._fn("ExtractMatrixColumn")(self.ih, byref(c_int(column_number + 1)), column_out.ih)

    def Scale(self, constant):
        self._fn("ScaleMatrix")(self.ih, _d(constant))

    def Increment(self, matB, alpha=1.0, threshold=0.0):
        """this = alpha*matB + this"""
        self._fn("IncrementMatrix")(matB.ih, self.ih, _d(alpha), _d(threshold))

    def Dot(self, matB):
        if self.is_complex:
            re, im = c_double(), c_double()
            self._fn("DotMatrix")(self.ih, matB.ih, byref(re), byref(im))
            return complex(re.value, im.value)
        v = c_double()
        self._fn("DotMatrix")(self.ih, matB.ih, byref(v))
        return v.value

    def PairwiseMultiply(self, matA, matB):
        self._fn("PairwiseMultiplyMatrix")(matA.ih, matB.ih, self.ih)

    def Gemm(self, matA, matB, isATransposed, isBTransposed, alpha, beta, threshold, memory_pool=None):
        """this = alpha*op(matA)*op(matB) + beta*this"""
        self._fn("MatrixMultiply")(matA.ih, matB.ih, self.ih, byref(c_bool(bool(isATransposed))),
                                   byref(c_bool(bool(isBTransposed))), _d(alpha), _d(beta), _d(threshold),
                                   memory_pool.ih if memory_pool is not None else None)

    def DiagonalScale(self, tlist):
        self._fn("MatrixDiagonalScale")(self.ih, tlist.ih)

    def Transpose(self, matA):
        self._fn("TransposeMatrix")(matA.ih, self.ih)

    def Print(self):
        self._fn("PrintMatrix")(self.ih)

    def WriteToMatrixMarket(self, file_name):
        b = file_name.encode()
        self._fn("PrintMatrixF")(self.ih, c_char_p(b), _i(len(b)))

    def MatrixToTripletList(self, triplet_list):
        self._fn("MatrixToTripletList")(self.ih, triplet_list.ih)

    # bulk host transfer (extension)
    @classmethod
    def from_scipy(cls, m):
        import scipy.sparse as sp
        m = sp.coo_matrix(m)
        tl = TripletList_c() if cls.is_complex else TripletList_r()
        tl.set_arrays(m.row + 1, m.col + 1, m.data)
        return cls(tl, m.shape[0], m.shape[1])

    def to_scipy(self):
        import scipy.sparse as sp
        tl = TripletList_c() if self.is_complex else TripletList_r()
        self.MatrixToTripletList(tl)
        rows, cols, vals = tl.get_arrays()
        return sp.coo_matrix((vals, (rows - 1, cols - 1)), shape=(self.GetRows(), self.GetColumns())).tocsc()


class Matrix_lsr(_Matrix_ls):
    _sfx = "lsr"
    is_complex = False


class Matrix_lsc(_Matrix_ls):
    _sfx = "lsc"
    is_complex = True

    def Conjugate(self):
        lib().ConjugateMatrix_lsc_wrp(self.ih)


# ---------------------------------------------------------------------------
class Matrix_ps:
    """Distributed sparse matrix living in GPU memory (reference Source/CPlusPlus/PSMatrix.h)."""

    def __init__(self, arg, grid: ProcessGrid | None = None, is_complex=False):
        self.ih = _handle()
        if isinstance(arg, Matrix_ps):
            lib().ConstructEmptyMatrix_ps_wrp(self.ih, _i(arg.GetActualDimension()))
            lib().CopyMatrix_ps_wrp(arg.ih, self.ih)
        elif isinstance(arg, str):
            b = arg.encode()
            if grid is None:
                lib().ConstructMatrixFromMatrixMarket_ps_wrp(self.ih, c_char_p(b), _i(len(b)))
            else:
                lib().ConstructMatrixFromMatrixMarketPG_ps_wrp(self.ih, c_char_p(b), _i(len(b)), grid.ih)
        elif grid is not None:
            lib().ConstructEmptyMatrixPG_ps_wrp(self.ih, _i(arg), grid.ih)
        elif is_complex:
            lib().ntb_ConstructEmptyMatrixComplex_ps(self.ih, _i(arg), _i(1))
        else:
            lib().ConstructEmptyMatrix_ps_wrp(self.ih, _i(arg))

    def __del__(self):
        try:
            lib().DestructMatrix_ps_wrp(self.ih)
        except Exception:
            pass

    # container
    def WriteToMatrixMarket(self, file_name):
        b = file_name.encode()
        lib().WriteMatrixToMatrixMarket_ps_wrp(self.ih, c_char_p(b), _i(len(b)))

    def FillFromTripletList(self, tl):
        if tl.is_complex:
            lib().FillMatrixFromTripletList_psc_wrp(self.ih, tl.ih)
        else:
            lib().FillMatrixFromTripletList_psr_wrp(self.ih, tl.ih)

    def FillDistributedPermutation(self, perm: Permutation, permuterows=True):
        lib().FillMatrixPermutation_ps_wrp(self.ih, perm.ih, byref(c_bool(bool(permuterows))))

    def FillIdentity(self):
        lib().FillMatrixIdentity_ps_wrp(self.ih)

    def FillDense(self):
        lib().FillMatrixDense_ps_wrp(self.ih)

    def Resize(self, new_size):
        lib().ResizeMatrix_ps_wrp(self.ih, _i(new_size))

    def GetMatrixBlock(self, tl, start_row, end_row, start_column, end_column):
        """0-based bounds, ends exclusive, like the reference C++ class (PSMatrix.cc:129-150 adds 1 to all four before
        the C call; distributed_includes/GetMatrixBlock.f90 tests start <= index < end)"""
        fn = lib().GetMatrixBlock_psc_wrp if tl.is_complex else lib().GetMatrixBlock_psr_wrp
        fn(self.ih, tl.ih, _i(start_row + 1), _i(end_row + 1), _i(start_column + 1), _i(end_column + 1))

    def GetMatrixSlice(self, submatrix, start_row, end_row, start_column, end_column):
        """0-based inclusive bounds like the reference C++ class (PSMatrix.cc adds 1 before the C call)"""
        lib().GetMatrixSlice_wrp(self.ih, submatrix.ih, _i(start_row + 1), _i(end_row + 1), _i(start_column + 1),
                                 _i(end_column + 1))

    def SnapToSparsityPattern(self, pattern):
        """reference MatrixConversion::SnapMatrixToSparsityPattern(mata, matb)"""
        lib().SnapMatrixToSparsityPattern_wrp(self.ih, pattern.ih)

    def GetActualDimension(self):
        v = c_int()
        lib().GetMatrixActualDimension_ps_wrp(self.ih, byref(v))
        return v.value

    def GetLogicalDimension(self):
        v = c_int()
        lib().GetMatrixLogicalDimension_ps_wrp(self.ih, byref(v))
        return v.value

    def GetSize(self):
        v = c_long()
        lib().GetMatrixSize_ps_wrp(self.ih, byref(v))
        return v.value

    def GetTripletList(self, tl):
        if tl.is_complex:
            lib().GetMatrixTripletList_psc_wrp(self.ih, tl.ih)
        else:
            lib().GetMatrixTripletList_psr_wrp(self.ih, tl.ih)

    def IsIdentity(self):
        return bool(lib().IsIdentity_ps_wrp(self.ih))

    def IsComplex(self):
        return bool(lib().ntb_MatrixIsComplex_ps(self.ih))

    def Transpose(self, matA):
        lib().TransposeMatrix_ps_wrp(matA.ih, self.ih)

    def Conjugate(self):
        lib().ConjugateMatrix_ps_wrp(self.ih)

    # algebra — the hot path
    def Dot(self, matB):
        v = c_double()
        lib().DotMatrix_psr_wrp(self.ih, matB.ih, byref(v))
        return v.value

    def Dot_c(self, matB):
        re, im = c_double(), c_double()
        lib().DotMatrix_psc_wrp(self.ih, matB.ih, byref(re), byref(im))
        return complex(re.value, im.value)

    def Increment(self, matB, alpha=1.0, threshold=0.0):
        """this = alpha*matB + this"""
        lib().IncrementMatrix_ps_wrp(matB.ih, self.ih, _d(alpha), _d(threshold))

    def PairwiseMultiply(self, matA, matB):
        lib().MatrixPairwiseMultiply_ps_wrp(matA.ih, matB.ih, self.ih)

    def Gemm(self, matA, matB, memory_pool: PMatrixMemoryPool | None = None, alpha=1.0, beta=0.0,
             threshold=0.0):
        """this = alpha*matA*matB + beta*this"""
        lib().MatrixMultiply_ps_wrp(matA.ih, matB.ih, self.ih, _d(alpha), _d(beta), _d(threshold),
                                    memory_pool.ih if memory_pool is not None else None)

    def GemmShift(self, matA, matB, identity, sigma, memory_pool: PMatrixMemoryPool | None = None, alpha=1.0,
                  threshold=0.0):
        """this = alpha*matA*matB (thresholded), then this += sigma*identity (the drivers' Gemm + Increment pair)"""
        lib().ntb_MatrixMultiplyShift_ps(matA.ih, matB.ih, self.ih, _d(alpha), _d(threshold), _d(sigma), identity.ih,
                                         memory_pool.ih if memory_pool is not None else None)

    def Scale(self, constant):
        if isinstance(constant, complex):
            lib().ntb_ScaleMatrixComplex_ps(self.ih, _d(constant.real), _d(constant.imag))
        else:
            lib().ScaleMatrix_ps_wrp(self.ih, _d(constant))

    def Norm(self):
        return lib().MatrixNorm_ps_wrp(self.ih)

    def MeasureAsymmetry(self):
        return lib().MeasureAsymmetry_ps_wrp(self.ih)

    def Trace(self):
        v = c_double()
        lib().MatrixTrace_ps_wrp(self.ih, byref(v))
        return v.value

    def Symmetrize(self):
        lib().SymmetrizeMatrix_ps_wrp(self.ih)

    def DiagonalScale(self, tlist):
        if tlist.is_complex:
            lib().MatrixDiagonalScale_psc_wrp(self.ih, tlist.ih)
        else:
            lib().MatrixDiagonalScale_psr_wrp(self.ih, tlist.ih)

    def Filter(self, threshold):
        lib().ntb_FilterMatrix_ps(self.ih, _d(threshold))

    # bulk host transfer (extension): global 1-based (row, col, value) arrays
    def fill_from_arrays(self, rows, cols, vals):
        rows = np.ascontiguousarray(rows, np.int32)
        cols = np.ascontiguousarray(cols, np.int32)
        cplx = np.iscomplexobj(vals)
        vals = np.ascontiguousarray(vals, np.complex128 if cplx else np.float64)
        lib().ntb_FillMatrixFromArrays_ps(self.ih, c_longlong(len(rows)), _ip(rows), _ip(cols),
                                          vals.view(np.float64).ctypes.data_as(POINTER(c_double)),
                                          c_int(1 if cplx else 0))

    def fill_from_staged(self, staged):
        """second half of the two-stage ingest: the matrix is built from a list staged with stage_arrays()"""
        lib().ntb_FillMatrixFromStaged_ps(self.ih, staged.ih)
        staged.keep = None

    def get_arrays(self, out=None):
        """local block as global 1-based triplets; `out` = (rows, cols, vals) buffers to fill (e.g. pinned memory,
        at least GetMatrixLocalSize entries each) instead of fresh arrays"""
        n = lib().ntb_GetMatrixLocalSize_ps(self.ih)
        if out is not None:
            rows, cols, vals = (a[:n] for a in out)
            assert rows.dtype == np.int32 and cols.dtype == np.int32
        else:
            rows, cols = np.empty(n, np.int32), np.empty(n, np.int32)
            vals = np.empty(n, np.complex128 if self.IsComplex() else np.float64)
        if n:
            lib().ntb_GetMatrixArrays_ps(self.ih, _ip(rows), _ip(cols),
                                         vals.view(np.float64).ctypes.data_as(POINTER(c_double)))
        return rows, cols, vals

    def get_arrays_async(self, out):
        """real matrices: like get_arrays(out=...) but the device-to-host copies run on a second stream; the (pinned)
        buffers are complete after egress_wait()"""
        cap = min(len(out[0]), len(out[1]), len(out[2]))
        n = int(lib().ntb_GetMatrixArraysAsync_ps(self.ih, c_longlong(cap), _ip(out[0]), _ip(out[1]),
                                                  out[2].ctypes.data_as(POINTER(c_double))))
        if n < 0:
            raise ValueError(f"get_arrays_async: the buffers hold {cap} entries, the local block has {-n}")
        return tuple(a[:n] for a in out)

    def fill_from_scipy(self, m):
        import scipy.sparse as sp
        m = sp.coo_matrix(m)
        self.fill_from_arrays(m.row + 1, m.col + 1, m.data)

    def to_scipy(self):
        """Local block of this rank as a scipy matrix of the ACTUAL dimension (single rank: the matrix)."""
        import scipy.sparse as sp
        rows, cols, vals = self.get_arrays()
        n = self.GetLogicalDimension()
        a = self.GetActualDimension()
        return sp.coo_matrix((vals, (rows - 1, cols - 1)), shape=(n, n)).tocsc()[:a, :a]

    def algorithmic_bytes(self):
        return lib().ntb_MatrixAlgorithmicBytes_ps(self.ih)


# ---------------------------------------------------------------------------
# solver classes (static methods, like the reference C++ classes)
# ---------------------------------------------------------------------------
def _density(fn, H, ISQ, trace, K, sp):
    e, mu = c_double(), c_double()
    fn(H.ih, ISQ.ih, _d(trace), K.ih, byref(e), byref(mu), sp.ih)
    return e.value, mu.value


class DensityMatrixSolvers:
    @staticmethod
    def TRS2(H, ISQ, trace, K, sp):
        return _density(lib().TRS2_wrp, H, ISQ, trace, K, sp)

    @staticmethod
    def TRS4(H, ISQ, trace, K, sp):
        return _density(lib().TRS4_wrp, H, ISQ, trace, K, sp)

    @staticmethod
    def PM(H, ISQ, trace, K, sp):
        return _density(lib().PM_wrp, H, ISQ, trace, K, sp)

    @staticmethod
    def HPCP(H, ISQ, trace, K, sp):
        return _density(lib().HPCP_wrp, H, ISQ, trace, K, sp)

    @staticmethod
    def ScaleAndFold(H, ISQ, trace, K, homo, lumo, sp):
        e = c_double()
        lib().ScaleAndFold_wrp(H.ih, ISQ.ih, _d(trace), K.ih, _d(homo), _d(lumo), byref(e), sp.ih)
        return e.value

    @staticmethod
    def EnergyDensityMatrix(H, D, ED, threshold=0.0):
        lib().EnergyDensityMatrix_wrp(H.ih, D.ih, ED.ih, _d(threshold))

    @staticmethod
    def McWeenyStep(D, DOut, S=None, threshold=0.0):
        if S is None:
            lib().McWeenyStep_wrp(D.ih, DOut.ih, _d(threshold))
        else:
            lib().McWeenyStepS_wrp(D.ih, DOut.ih, S.ih, _d(threshold))


class SignSolvers:
    @staticmethod
    def ComputeSign(M, Out, sp):
        lib().SignFunction_wrp(M.ih, Out.ih, sp.ih)

    @staticmethod
    def ComputePolarDecomposition(M, U, Hm, sp):
        lib().PolarDecomposition_wrp(M.ih, U.ih, Hm.ih if Hm is not None else None, sp.ih)


class InverseSolvers:
    @staticmethod
    def Invert(M, Out, sp):
        lib().Invert_wrp(M.ih, Out.ih, sp.ih)

    @staticmethod
    def PseudoInverse(M, Out, sp):
        lib().PseudoInverse_wrp(M.ih, Out.ih, sp.ih)


class SquareRootSolvers:
    @staticmethod
    def SquareRoot(M, Out, sp, order=None):
        if order is None:
            lib().SquareRoot_wrp(M.ih, Out.ih, sp.ih)
        else:
            lib().ntb_SquareRootOrder_wrp(M.ih, Out.ih, sp.ih, _i(order))

    @staticmethod
    def InverseSquareRoot(M, Out, sp, order=None):
        if order is None:
            lib().InverseSquareRoot_wrp(M.ih, Out.ih, sp.ih)
        else:
            lib().ntb_InverseSquareRootOrder_wrp(M.ih, Out.ih, sp.ih, _i(order))


class ExponentialSolvers:
    @staticmethod
    def ComputeExponential(M, Out, sp):
        lib().ComputeExponential_wrp(M.ih, Out.ih, sp.ih)


class EigenBounds:
    @staticmethod
    def GershgorinBounds(M):
        mn, mx = c_double(), c_double()
        lib().GershgorinBounds_wrp(M.ih, byref(mn), byref(mx))
        return mn.value, mx.value

    @staticmethod
    def PowerBounds(M, sp):
        v = c_double()
        lib().PowerBounds_wrp(M.ih, byref(v), sp.ih)
        return v.value


class MatrixConversion:
    @staticmethod
    def SnapMatrixToSparsityPattern(mata, matb):
        lib().SnapMatrixToSparsityPattern_wrp(mata.ih, matb.ih)


class LoadBalancer:
    @staticmethod
    def PermuteMatrix(mat_in, mat_out, perm, pool=None):
        lib().PermuteMatrix_wrp(mat_in.ih, mat_out.ih, perm.ih, pool.ih if pool is not None else None)

    @staticmethod
    def UndoPermuteMatrix(mat_in, mat_out, perm, pool=None):
        lib().UndoPermuteMatrix_wrp(mat_in.ih, mat_out.ih, perm.ih, pool.ih if pool is not None else None)


# ---------------------------------------------------------------------------
# counters (extension)
# ---------------------------------------------------------------------------
def counters():
    out = (c_double * 4)()
    lib().ntb_get_counters(out)
    return {"launches": int(out[0]), "multiplies": int(out[1]), "flops": float(out[2]),
            "dense_rule_blocks": int(out[3])}


def reset_counters():
    lib().ntb_reset_counters()


def set_tile_path(on=True):
    lib().ntb_set_tile_path(c_int(1 if on else 0))


def tile_counters():
    out = (c_double * 2)()
    lib().ntb_get_tile_counters(out)
    return {"tile_products": int(out[0]), "dmma": float(out[1])}


def tile_builds():
    return int(lib().ntb_tile_builds())


def halo_counters():
    out = (c_double * 2)()
    lib().ntb_get_halo_counters(out)
    return {"products": int(out[0]), "bytes": float(out[1])}


def peer_counters():
    out = (c_double * 4)()
    lib().ntb_get_peer_counters(out)
    return {"ok": bool(out[0]), "products": int(out[1]), "exchanges": int(out[2]), "slab_peak_bytes": int(out[3])}


def sync_count():
    """host waits for the library stream since the last reset_counters()"""
    return int(lib().ntb_get_sync_count())


def measure_dmma_peak_tflops(repeats=3):
    """issue peak of DMMA.8x8x4 on this GPU (TFLOP/s), measured now"""
    return float(lib().ntb_measure_dmma_peak_tflops(c_int(repeats)))


def set_halo_path(on=True):
    lib().ntb_set_halo_path(c_int(1 if on else 0))


def tile_combine(P, Q, Out, mode=0, alpha=1.0, beta=1.0, threshold=0.0, sigma=0.0):
    """tile-space combination (csrc/spgemm_tile.cu: tile_combine); False: the operands do not live as tile forms"""
    return bool(lib().ntb_TileCombine_ps(P.ih, Q.ih, c_int(mode), c_double(alpha), c_double(beta), c_double(threshold),
                                         c_double(sigma), Out.ih))


def tile_scalars(mode, A, B=None):
    out = (c_double * 2)()
    ok = lib().ntb_TileScalars_ps(c_int(mode), A.ih, B.ih if B is not None else None, out)
    return (float(out[0]), float(out[1])) if ok else None


def set_fused_steps(on=True):
    lib().ntb_set_fused_steps(c_int(1 if on else 0))


def hash_columns():
    return int(lib().ntb_hash_columns())


def complex_tile_products():
    return int(lib().ntb_complex_tile_products())


def tile_combines():
    return int(lib().ntb_tile_combines())


def set_permute_gemm(on=True):
    lib().ntb_set_permute_gemm(c_int(1 if on else 0))


def set_fused_shift(on=True):
    lib().ntb_set_fused_shift(c_int(1 if on else 0))


def set_fused_norm(on=True):
    lib().ntb_set_fused_norm(c_int(1 if on else 0))


def fused_norms():
    return int(lib().ntb_fused_norms())


def sign_iteration(X, identity, T1, T2, alpha_k, threshold, memory_pool=None):
    """one pass of the SignFunction driver's loop body; X advances in place, returns ||X_new - X_old||"""
    return float(lib().ntb_SignIteration(X.ih, identity.ih, T1.ih, T2.ih, _d(alpha_k), _d(threshold),
                                         memory_pool.ih if memory_pool is not None else None))


def sign_step(X, identity, T1, Xnext, alpha_k, threshold, memory_pool=None):
    """the same loop body out of place: Xnext <- next iterate, X untouched; returns ||Xnext - X||"""
    return float(lib().ntb_SignStep(X.ih, identity.ih, T1.ih, Xnext.ih, _d(alpha_k), _d(threshold),
                                    memory_pool.ih if memory_pool is not None else None))


def set_flop_counting(on=True):
    """instrumentation (off by default): count the useful products of every multiply"""
    lib().ntb_set_flop_counting(c_int(1 if on else 0))


def deferred_counters():
    out = (c_double * 2)()
    lib().ntb_get_deferred_counters(out)
    return {"products": int(out[0]), "materialized": int(out[1])}


class StagedArrays:
    """a real global 1-based (rows, cols, vals) list whose host-to-device copies are in flight on the library's copy
    stream (first half of the two-stage ingest); the host arrays should be pinned and are kept alive here"""

    def __init__(self, rows, cols, vals):
        assert rows.dtype == np.int32 and cols.dtype == np.int32 and vals.dtype == np.float64
        self.keep = (rows, cols, vals)
        self.ih = _handle()
        lib().ntb_StageArrays(self.ih, c_longlong(len(rows)), _ip(rows), _ip(cols), _dp(vals))


def stage_arrays(rows, cols, vals):
    return StagedArrays(rows, cols, vals)


def sorted_ingests():
    return int(lib().ntb_sorted_ingests())


def egress_wait():
    lib().ntb_EgressWait()


def algorithmic_bytes():
    return float(lib().ntb_algorithmic_bytes())


def profile_enable(on=True):
    lib().ntb_profile_enable(c_int(1 if on else 0))


def profile_read():
    out = (c_double * 2)()
    lib().ntb_profile_read(out)
    return {"numeric_ms": float(out[0]), "products": int(out[1])}


def profile_read_phases():
    out = (c_double * 8)()
    lib().ntb_profile_read_phases(out)
    return {"symbolic_ms": float(out[0]), "tail_ms": float(out[2]), "scalars_ms": float(out[3]), "combine_ms": float(out[4]),
            "panel_gather_ms": float(out[5]), "operand_tile_build_ms": float(out[6]), "slice_sum_ms": float(out[7])}


def last_solve():
    out = (c_double * 5)()
    lib().ntb_last_solve(out)
    return {"loop_counter": int(out[0]), "last_value": float(out[1]), "energy": float(out[2]),
            "multiplies": int(out[3]), "flops": float(out[4])}


def grid_layout(rank, size, rows, cols, slices, matrix_dim):
    """Pure host arithmetic (no GPU needed): ownership of `rank` on a rows x cols x slices grid."""
    out = (c_int * 12)()
    lib().ntb_grid_layout(c_int(rank), c_int(size), c_int(rows), c_int(cols), c_int(slices), c_int(matrix_dim), out)
    keys = ["my_slice", "my_row", "my_col", "logical_dim", "local_rows", "local_cols", "start_row", "start_col",
            "row_blocks", "col_blocks", "row_comm_colour", "col_comm_colour"]
    return dict(zip(keys, list(out)))


def default_grid(size):
    out = (c_int * 3)()
    lib().ntb_default_grid(c_int(size), out)
    return tuple(out)


def set_stream(cuda_stream_ptr: int):
    lib().ntb_set_stream(c_void_p(cuda_stream_ptr))


def synchronize():
    lib().ntb_synchronize()
