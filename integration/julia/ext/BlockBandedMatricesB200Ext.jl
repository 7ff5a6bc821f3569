# BlockBandedMatricesB200Ext — package extension that routes the reference's MulAdd / Ldiv lazy ops
# to libbbm_b200.so when the matrix `data` lives in B200 memory.
#
# NOT EXECUTED IN THE BUILD IMAGE (no julia binary there): this file is the reference-side binding a
# maintainer would add; every ccall below names an entry point declared in include/bbm_b200.h and
# exercised through the same C ABI by the Python/ctypes harness of this repository.
#
# Activation, same mechanism as ext/BlockBandedMatricesSparseArraysExt.jl (Project.toml:13-17):
#
#   [weakdeps]
#   BBMB200 = "<uuid of a 20-line stub package that only defines `const libbbm = \"/path/libbbm_b200.so\"`>"
#   [extensions]
#   BlockBandedMatricesB200Ext = "BBMB200"
#
# What stays host-side Julia (unchanged): the BlockBandedMatrix / BandedBlockBandedMatrix
# constructors, block axes, blockbandwidths / subblockbandwidths, the MemoryLayout traits
# (src/AbstractBlockBandedMatrix.jl:8-30, src/BandedBlockBandedMatrix.jl:263-269,
# src/BlockSkylineMatrix.jl:369) and the MulAdd / Ldiv / materialize! dispatch of ArrayLayouts.
# What moves: `data` (shape unchanged) and the arithmetic.
module BlockBandedMatricesB200Ext

using BlockBandedMatrices
using BlockBandedMatrices: BandedBlockBandedMatrix, BlockSkylineMatrix, BandedBlockBandedColumns,
                           BlockBandedColumns, blockbandwidths, subblockbandwidths, colblockbandwidths
using BlockBandedMatrices.ArrayLayouts: MulAdd, Ldiv, TriangularLayout, MemoryLayout, DenseColumnMajor, AbstractColumnMajor
using LinearAlgebra: UniformScaling
import BlockBandedMatrices.ArrayLayouts: materialize!, triangulardata
using BlockBandedMatrices.BlockArrays: BlockedMatrix, blocklengths, blocksize
using BlockBandedMatrices.BandedMatrices: BandError
import BlockBandedMatrices.MatrixFactorizations
using LinearAlgebra
import BBMB200: libbbm

# ---- status codes (bbm_status, include/bbm_b200.h) -------------------------------------------
const BBM_OK, BBM_ERR_DIM, BBM_ERR_BAND, BBM_ERR_UNSUPPORTED = Int32(0), Int32(2), Int32(3), Int32(8)

mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(device::Integer=0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:bbm_create, libbbm), Int32, (Int32, Ref{Ptr{Cvoid}}), device, r)
        rc == 0 || error("bbm_create failed: ", unsafe_string(ccall((:bbm_status_string, libbbm), Cstring, (Int32,), rc)),
                         " (no CPU fallback: a B200 is required)")
        h = new(r[])
        finalizer(x -> ccall((:bbm_destroy, libbbm), Int32, (Ptr{Cvoid},), x.ptr), h)
    end
end
const default_handle = Ref{Handle}()
handle() = isassigned(default_handle) ? default_handle[] : (default_handle[] = Handle(0))

function check(rc::Int32, h::Handle=handle())
    rc == BBM_OK && return nothing
    msg = unsafe_string(ccall((:bbm_last_error, libbbm), Cstring, (Ptr{Cvoid},), h.ptr))
    rc == BBM_ERR_DIM  && throw(DimensionMismatch(msg))          # where [upstream] BlockArrays / src/linalg.jl:42,61 throw
    rc == BBM_ERR_BAND && throw(BandError(nothing, 0))           # src/BlockSkylineMatrix.jl:491
    error("libbbm_b200 [status $rc]: $msg")
end

# ---- device buffers: the storage type the extension's methods are keyed on ------------------
"Dense column-major array in B200 memory (owned through bbm_malloc / bbm_free)."
mutable struct B200Array{T,N} <: DenseArray{T,N}
    ptr::Ptr{T}
    dims::NTuple{N,Int}
    # descriptors of the matrix structures laid over this buffer (structure key => (kind, bbm_*_desc_t)); they live and
    # die with the buffer -- the containers themselves are immutable structs and cannot carry finalizers
    descs::Dict{Any,Tuple{Symbol,Ptr{Cvoid}}}
    function B200Array{T,N}(::UndefInitializer, dims::NTuple{N,Int}) where {T,N}
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:bbm_malloc, libbbm), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Csize_t), handle().ptr, r, sizeof(T) * prod(dims)))
        a = new{T,N}(Ptr{T}(r[]), dims, Dict{Any,Tuple{Symbol,Ptr{Cvoid}}}())
        finalizer(a) do x
            for (kind, p) in values(x.descs)
                kind === :bbb ? ccall((:bbm_bbb_desc_destroy, libbbm), Int32, (Ptr{Cvoid},), p) :
                                ccall((:bbm_sky_desc_destroy, libbbm), Int32, (Ptr{Cvoid},), p)
            end
            ccall((:bbm_free, libbbm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), handle().ptr, x.ptr)
        end
    end
end
B200Array{T}(::UndefInitializer, dims::Integer...) where T = B200Array{T,length(dims)}(undef, map(Int, dims))
B200Array{T}(::UndefInitializer, dims::NTuple{N,Integer}) where {T,N} = B200Array{T,N}(undef, map(Int, dims))
const B200Vector{T} = B200Array{T,1}
const B200Matrix{T} = B200Array{T,2}
Base.size(a::B200Array) = a.dims
Base.similar(a::B200Array{T}, ::Type{S}, dims::Dims{N}) where {T,S,N} = B200Array{S,N}(undef, dims)
Base.unsafe_convert(::Type{Ptr{T}}, a::B200Array{T}) where T = a.ptr
Base.strides(a::B200Array) = Base.size_to_strides(1, a.dims...)
MemoryLayout(::Type{<:B200Array}) = DenseColumnMajor()
Base.getindex(a::B200Array, i...) = error("scalar indexing of device memory; copy to the host with Array(a)")

"One memcpy of the parent array, shape unchanged."
function B200Array(a::Array{T,N}) where {T,N}
    d = B200Array{T,N}(undef, size(a))
    check(ccall((:bbm_memcpy_h2d, libbbm), Int32, (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Csize_t), handle().ptr, d.ptr, a, sizeof(a)))
    check(ccall((:bbm_synchronize, libbbm), Int32, (Ptr{Cvoid},), handle().ptr))
    d
end
function Base.Array(d::B200Array{T,N}) where {T,N}
    a = Array{T,N}(undef, size(d))
    check(ccall((:bbm_memcpy_d2h, libbbm), Int32, (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Csize_t), handle().ptr, a, d.ptr, sizeof(a)))
    check(ccall((:bbm_synchronize, libbbm), Int32, (Ptr{Cvoid},), handle().ptr))
    a
end

"Device-resident copy of a BandedBlockBandedMatrix: same axes and bandwidths, `data` parent in HBM."
function b200(A::BandedBlockBandedMatrix{T}) where T<:Union{Float32,Float64}
    l, u = blockbandwidths(A); λ, μ = subblockbandwidths(A)
    data = BlockedMatrix(B200Array(Matrix(parent(A.data))), axes(A.data))
    BlockBandedMatrices._BandedBlockBandedMatrix(data, A.raxis, (l, u), (λ, μ))
end
"Device-resident copy of a BlockSkylineMatrix / BlockBandedMatrix: same BlockSkylineSizes, `data` vector in HBM."
b200(A::BlockSkylineMatrix{T}) where T<:Union{Float32,Float64} =
    BlockBandedMatrices._BlockSkylineMatrix(B200Array(Vector(A.data)), A.block_sizes)

const B200BBB{T} = BandedBlockBandedMatrix{T,<:BlockedMatrix{T,<:B200Matrix{T}}}
const B200Sky{T} = BlockSkylineMatrix{T,<:B200Vector{T}}

# ---- descriptors ------------------------------------------------------------------------------
# A descriptor depends on the block structure only and is cached in the device buffer it is used with (no global table:
# nothing keeps a dropped matrix alive, and its descriptors are destroyed by the buffer's finalizer).  Sub-views over the
# same buffer get their own entry (different structure key).
desc_table(buf::B200Array) = buf.descs

function descriptor(A::B200BBB)
    r = Int64.(blocklengths(axes(A, 1))); c = Int64.(blocklengths(axes(A, 2)))
    l, u = blockbandwidths(A); λ, μ = subblockbandwidths(A)
    tab = desc_table(parent(A.data))
    last(get!(tab, (:bbb, r, c, l, u, λ, μ)) do
        d = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:bbm_bbb_desc_create, libbbm), Int32,
                    (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Int64, Int64, Int64, Ref{Ptr{Cvoid}}),
                    handle().ptr, length(r), length(c), r, c, l, u, λ, μ, d))
        (:bbb, d[])
    end)
end
function descriptor(A::B200Sky)
    r = Int64.(blocklengths(axes(A, 1))); c = Int64.(blocklengths(axes(A, 2)))
    l, u = colblockbandwidths(A)
    lv, uv = Vector{Int64}(l), Vector{Int64}(u)
    tab = desc_table(A.data)
    last(get!(tab, (:sky, r, c, lv, uv)) do
        d = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:bbm_sky_desc_create, libbbm), Int32,
                    (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ref{Ptr{Cvoid}}),
                    handle().ptr, length(r), length(c), r, c, lv, uv, d))
        (:sky, d[])
    end)
end

suffix(::Type{Float64}) = :f64
suffix(::Type{Float32}) = :f32
ld(x::AbstractVecOrMat) = max(1, size(x, 1))

# ---- mul!(y, A, x, α, β)  ==  materialize!(MulAdd(α, A, x, β, y)) ----------------------------
# Replaces [upstream BlockArrays] materialize!(::MatMulVecAdd / ::MatMulMatAdd{<:AbstractBlockLayout,...})
# and its one-BLAS-gbmv-per-block loop (SURVEY.md 3.1).  Field names M.α, M.A, M.B, M.β, M.C as in
# src/linalg.jl:33,51.
for (T, suf) in ((Float64, "f64"), (Float32, "f32"))
    @eval function materialize!(M::MulAdd{<:BandedBlockBandedColumns,<:Any,<:Any,$T,<:B200BBB{$T},<:B200Array{$T},<:B200Array{$T}})
        A, x, y = M.A, M.B, M.C
        size(A, 2) == size(x, 1) && size(A, 1) == size(y, 1) && size(x, 2) == size(y, 2) ||
            throw(DimensionMismatch("A has size $(size(A)), x has size $(size(x)), y has size $(size(y))"))
        check(ccall(($(QuoteNode(Symbol("bbm_bbb_mul_", suf))), libbbm), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, $T, Ptr{$T}, Ptr{$T}, Int64, Int64, $T, Ptr{$T}, Int64),
                    handle().ptr, descriptor(A), M.α, parent(A.data).ptr, x.ptr, ld(x), size(x, 2), M.β, y.ptr, ld(y)))
        y
    end
    # host vectors with a device-resident matrix: one call that uploads x, multiplies and downloads y
    @eval function materialize!(M::MulAdd{<:BandedBlockBandedColumns,<:Any,<:Any,$T,<:B200BBB{$T},<:StridedVecOrMat{$T},<:StridedVecOrMat{$T}})
        A, x, y = M.A, M.B, M.C
        size(A, 2) == size(x, 1) && size(A, 1) == size(y, 1) && size(x, 2) == size(y, 2) ||
            throw(DimensionMismatch("A has size $(size(A)), x has size $(size(x)), y has size $(size(y))"))
        check(ccall(($(QuoteNode(Symbol("bbm_bbb_mul_host_", suf))), libbbm), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, $T, Ptr{$T}, Ptr{$T}, Int64, Int64, $T, Ptr{$T}, Int64),
                    handle().ptr, descriptor(A), M.α, parent(A.data).ptr, x, stride(x, 2), size(x, 2), M.β, y, stride(y, 2)))
        y
    end
    @eval function materialize!(M::MulAdd{<:BlockBandedColumns,<:Any,<:Any,$T,<:B200Sky{$T},<:B200Array{$T},<:B200Array{$T}})
        A, x, y = M.A, M.B, M.C
        size(A, 2) == size(x, 1) && size(A, 1) == size(y, 1) && size(x, 2) == size(y, 2) ||
            throw(DimensionMismatch("A has size $(size(A)), x has size $(size(x)), y has size $(size(y))"))
        check(ccall(($(QuoteNode(Symbol("bbm_sky_mul_", suf))), libbbm), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, $T, Ptr{$T}, Ptr{$T}, Int64, Int64, $T, Ptr{$T}, Int64),
                    handle().ptr, descriptor(A), M.α, A.data.ptr, x.ptr, ld(x), size(x, 2), M.β, y.ptr, ld(y)))
        y
    end
    # Y <- α X A + β Y for a dense device matrix X: `X*A`, `similar(X) .= MulAdd(X, A)` (test/test_linalg.jl:56-57)
    @eval function materialize!(M::MulAdd{<:AbstractColumnMajor,<:BlockBandedColumns,<:AbstractColumnMajor,$T,<:B200Array{$T,2},<:B200Sky{$T},<:B200Array{$T,2}})
        X, A, Y = M.A, M.B, M.C
        size(X, 2) == size(A, 1) && size(X, 1) == size(Y, 1) && size(A, 2) == size(Y, 2) ||
            throw(DimensionMismatch("X has size $(size(X)), A has size $(size(A)), Y has size $(size(Y))"))
        check(ccall(($(QuoteNode(Symbol("bbm_sky_rmul_", suf))), libbbm), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, $T, Ptr{$T}, Int64, Int64, Ptr{$T}, $T, Ptr{$T}, Int64),
                    handle().ptr, descriptor(A), M.α, X.ptr, ld(X), size(X, 1), A.data.ptr, M.β, Y.ptr, ld(Y)))
        Y
    end
    # C <- α A B + β C, all block-skyline (result structure from similar(::MulAdd), src/linalg.jl:32-48,
    # which returns a B200-backed BlockSkylineMatrix because `similar(::B200Vector)` is a B200Vector)
    @eval function materialize!(M::MulAdd{<:BlockBandedColumns,<:BlockBandedColumns,<:BlockBandedColumns,$T,<:B200Sky{$T},<:B200Sky{$T},<:B200Sky{$T}})
        check(ccall(($(QuoteNode(Symbol("bbm_sky_matmul_", suf))), libbbm), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, $T, Ptr{$T}, Ptr{$T}, $T, Ptr{$T}),
                    handle().ptr, descriptor(M.A), descriptor(M.B), descriptor(M.C), M.α, M.A.data.ptr, M.B.data.ptr, M.β, M.C.data.ptr))
        M.C
    end
    # mul!(y, UpperTriangular(A), x) etc. for a triangular BandedBlockBandedMatrix (test/test_triblockbanded.jl:17-53)
    @eval function materialize!(M::MulAdd{<:TriangularLayout{UPLO,UNIT,<:BandedBlockBandedColumns},<:Any,<:Any,$T,<:AbstractTriangular{$T,<:B200BBB{$T}},<:B200Array{$T},<:B200Array{$T}}) where {UPLO,UNIT}
        A, x, y = triangulardata(M.A), M.B, M.C
        size(A, 2) == size(x, 1) && size(A, 1) == size(y, 1) && size(x, 2) == size(y, 2) ||
            throw(DimensionMismatch("A has size $(size(A)), x has size $(size(x)), y has size $(size(y))"))
        rc = ccall(($(QuoteNode(Symbol("bbm_bbb_trimul_", suf))), libbbm), Int32,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, $T, Ptr{$T}, Ptr{$T}, Int64, Int64, $T, Ptr{$T}, Int64),
                   handle().ptr, descriptor(A), Int32(UPLO), Int32(UNIT == 'U'), M.α, parent(A.data).ptr, x.ptr, ld(x), size(x, 2), M.β, y.ptr, ld(y))
        rc == BBM_ERR_UNSUPPORTED && return invoke(materialize!, Tuple{MulAdd}, M)   # blocks that do not match: the reference's generic path
        check(rc)
        y
    end
    # mul!(y, A', x) / transpose(A)*x for real T: the adjoint's Rows layout (src/adjtransblockbanded.jl:2-7)
    @eval function materialize!(M::MulAdd{<:BlockBandedMatrices.BandedBlockBandedRows,<:Any,<:Any,$T,<:Union{Adjoint{$T,<:B200BBB{$T}},Transpose{$T,<:B200BBB{$T}}},<:B200Array{$T},<:B200Array{$T}})
        A, x, y = parent(M.A), M.B, M.C
        size(A, 1) == size(x, 1) && size(A, 2) == size(y, 1) && size(x, 2) == size(y, 2) ||
            throw(DimensionMismatch("A' has size $(reverse(size(A))), x has size $(size(x)), y has size $(size(y))"))
        check(ccall(($(QuoteNode(Symbol("bbm_bbb_mul_t_", suf))), libbbm), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, $T, Ptr{$T}, Ptr{$T}, Int64, Int64, $T, Ptr{$T}, Int64),
                    handle().ptr, descriptor(A), M.α, parent(A.data).ptr, x.ptr, ld(x), size(x, 2), M.β, y.ptr, ld(y)))
        y
    end
    # C <- α A B + β C, all banded-block-banded (result structure from similar(::MulAdd), src/linalg.jl:50-71)
    @eval function materialize!(M::MulAdd{<:BandedBlockBandedColumns,<:BandedBlockBandedColumns,<:BandedBlockBandedColumns,$T,<:B200BBB{$T},<:B200BBB{$T},<:B200BBB{$T}})
        check(ccall(($(QuoteNode(Symbol("bbm_bbb_matmul_", suf))), libbbm), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, $T, Ptr{$T}, Ptr{$T}, $T, Ptr{$T}),
                    handle().ptr, descriptor(M.A), descriptor(M.B), descriptor(M.C), M.α, parent(M.A.data).ptr, parent(M.B.data).ptr, M.β, parent(M.C.data).ptr))
        M.C
    end
    # axpy!(a, X, Y) on identical structures and lmul!(a, A): level-1 kernels on `data`
    # (src/BandedBlockBandedMatrix.jl:605-612)
    @eval function LinearAlgebra.axpy!(a::$T, X::B200BBB{$T}, Y::B200BBB{$T})
        (axes(X) == axes(Y) && blockbandwidths(X) == blockbandwidths(Y) && subblockbandwidths(X) == subblockbandwidths(Y)) ||
            return invoke(LinearAlgebra.axpy!, Tuple{$T,BandedBlockBandedMatrix{$T},BandedBlockBandedMatrix{$T}}, a, X, Y)
        check(ccall(($(QuoteNode(Symbol("bbm_axpy_", suf))), libbbm), Int32, (Ptr{Cvoid}, Int64, $T, Ptr{$T}, Ptr{$T}),
                    handle().ptr, length(parent(X.data)), a, parent(X.data).ptr, parent(Y.data).ptr))
        Y
    end
    @eval function LinearAlgebra.lmul!(a::$T, A::B200BBB{$T})
        check(ccall(($(QuoteNode(Symbol("bbm_scal_", suf))), libbbm), Int32, (Ptr{Cvoid}, Int64, $T, Ptr{$T}),
                    handle().ptr, length(parent(A.data)), a, parent(A.data).ptr))
        A
    end
    # ldiv!(UpperTriangular(A), B) etc.  ==  materialize!(Ldiv(U, B)); callers: src/blockskylineqr.jl:136-155
    @eval function materialize!(L::Ldiv{<:TriangularLayout{UPLO,UNIT,<:BlockBandedColumns},<:Any,<:AbstractTriangular{$T,<:B200Sky{$T}},<:B200Array{$T}}) where {UPLO,UNIT}
        A, B = triangulardata(L.A), L.B
        size(A, 1) == size(B, 1) || throw(DimensionMismatch("T has size $(size(A)), B has size $(size(B))"))
        check(ccall(($(QuoteNode(Symbol("bbm_sky_trsm_", suf))), libbbm), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Ptr{$T}, Ptr{$T}, Int64, Int64),
                    handle().ptr, descriptor(A), Int32(UPLO), Int32(UNIT == 'U'), A.data.ptr, B.ptr, ld(B), size(B, 2)))
        B
    end
    # ldiv!(LowerTriangular(A), b) etc. for a triangular BandedBlockBandedMatrix (test/test_triblockbanded.jl:80-99,
    # the x .= Ldiv(L, x) of examples/finitedifference_2d.jl:21); BBM_ERR_UNSUPPORTED (blocks that do not match)
    # falls through to the reference's generic solve
    @eval function materialize!(L::Ldiv{<:TriangularLayout{UPLO,UNIT,<:BandedBlockBandedColumns},<:Any,<:AbstractTriangular{$T,<:B200BBB{$T}},<:B200Array{$T}}) where {UPLO,UNIT}
        A, B = triangulardata(L.A), L.B
        size(A, 1) == size(B, 1) || throw(DimensionMismatch("T has size $(size(A)), B has size $(size(B))"))
        rc = ccall(($(QuoteNode(Symbol("bbm_bbb_trsv_", suf))), libbbm), Int32,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Ptr{$T}, Ptr{$T}, Int64, Int64),
                   handle().ptr, descriptor(A), Int32(UPLO), Int32(UNIT == 'U'), parent(A.data).ptr, B.ptr, ld(B), size(B, 2))
        rc == BBM_ERR_UNSUPPORTED && return invoke(materialize!, Tuple{Ldiv}, L)
        check(rc)
        B
    end
end

# sparse(A) built on the device (ext/BlockBandedMatricesSparseArraysExt.jl:10-27); needs SparseArrays loaded by the caller
function b200_sparse_arrays(A::B200BBB{T}) where {T<:Union{Float32,Float64}}
    colptr = B200Array{Int64}(undef, size(A, 2) + 1)
    nnz = Ref{Int64}(0)
    check(ccall((:bbm_bbb_sparse_colptr, libbbm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int64}, Ref{Int64}),
                handle().ptr, descriptor(A), Int32(1), colptr.ptr, nnz))
    rowval, nzval = B200Array{Int64}(undef, nnz[]), B200Array{T}(undef, nnz[])
    fill = T === Float64 ? :bbm_bbb_sparse_fill_f64 : :bbm_bbb_sparse_fill_f32
    check(ccall((fill, libbbm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{T}, Ptr{Int64}, Int32, Ptr{Int64}, Ptr{T}),
                handle().ptr, descriptor(A), parent(A.data).ptr, colptr.ptr, Int32(1), rowval.ptr, nzval.ptr))
    Array(colptr), Array(rowval), Array(nzval)      # SparseMatrixCSC(size(A)..., colptr, rowval, nzval)
end

# ldiv!(F, B) / lmul!(F.Q', B) for F = qr(A) computed by the reference on the host and moved with b200(F):
# src/blockskylineqr.jl:77-127 (Q'B panel by panel) and :131-145 (then the triangular solve).  Float64 and Float32
# (the Float32 entry points run the Float64 kernels on widened copies and round once at the end).
b200(F::MatrixFactorizations.QR{T,<:BlockSkylineMatrix}) where T<:Union{Float32,Float64} = MatrixFactorizations.QR(b200(F.factors), B200Array(F.τ))
for (T, suf) in ((Float64, "f64"), (Float32, "f32"))
@eval begin
# qr!(A) on the device (src/blockskylineqr.jl:45-49); `_qr` (:75) has already widened A to (l, l+u) with similar/copyto!
function BlockBandedMatrices.qr!(A::B200Sky{$T})
    M, N = blocksize(A)
    τ = B200Array{$T}(undef, M < N ? size(A, 1) : size(A, 2))
    check(ccall(($(QuoteNode(Symbol("bbm_sky_qr_", suf))), libbbm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{$T}, Ptr{$T}),
                handle().ptr, descriptor(A), A.data.ptr, τ.ptr))
    MatrixFactorizations.QR(A, τ)
end
function LinearAlgebra.ldiv!(F::MatrixFactorizations.QR{$T,<:B200Sky{$T}}, B::B200Array{$T})
    A = F.factors
    size(A, 1) == size(B, 1) || throw(DimensionMismatch(string("F has size ", size(A), ", B has size ", size(B))))
    rc = ccall(($(QuoteNode(Symbol("bbm_sky_qr_ldiv_", suf))), libbbm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{$T}, Ptr{$T}, Ptr{$T}, Int64, Int64),
               handle().ptr, descriptor(A), A.data.ptr, F.τ.ptr, B.ptr, ld(B), size(B, 2))
    rc == BBM_ERR_UNSUPPORTED && error("thin/wide block-banded QR solves stay on the host: use Array(B)")
    check(rc)
    B
end
# the QL counterparts (src/blockskylineqr.jl:51-72, 96-111, 149-157); `_ql` (:76) has widened A to (l+u, u)
function BlockBandedMatrices.ql!(A::B200Sky{$T})
    M, N = blocksize(A)
    M < N && throw(ArgumentError("Wide block-QL not implented"))
    τ = B200Array{$T}(undef, size(A, 2))
    check(ccall(($(QuoteNode(Symbol("bbm_sky_ql_", suf))), libbbm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{$T}, Ptr{$T}),
                handle().ptr, descriptor(A), A.data.ptr, τ.ptr))
    MatrixFactorizations.QL(A, τ)
end
function LinearAlgebra.ldiv!(F::MatrixFactorizations.QL{$T,<:B200Sky{$T}}, B::B200Array{$T})
    A = F.factors
    size(A, 1) == size(B, 1) || throw(DimensionMismatch(string("F has size ", size(A), ", B has size ", size(B))))
    check(ccall(($(QuoteNode(Symbol("bbm_sky_ql_ldiv_", suf))), libbbm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{$T}, Ptr{$T}, Ptr{$T}, Int64, Int64),
                handle().ptr, descriptor(A), A.data.ptr, F.τ.ptr, B.ptr, ld(B), size(B, 2)))
    B
end
function LinearAlgebra.lmul!(adjQ::Adjoint{$T,<:MatrixFactorizations.QLPackedQ{$T,<:B200Sky{$T}}}, B::B200Array{$T})
    Q = parent(adjQ)
    check(ccall(($(QuoteNode(Symbol("bbm_sky_ql_lmul_adj_", suf))), libbbm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{$T}, Ptr{$T}, Ptr{$T}, Int64, Int64),
                handle().ptr, descriptor(Q.factors), Q.factors.data.ptr, Q.τ.ptr, B.ptr, ld(B), size(B, 2)))
    B
end
function LinearAlgebra.lmul!(adjQ::Adjoint{$T,<:MatrixFactorizations.QRPackedQ{$T,<:B200Sky{$T}}}, B::B200Array{$T})
    Q = parent(adjQ)
    check(ccall(($(QuoteNode(Symbol("bbm_sky_qr_lmul_adj_", suf))), libbbm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{$T}, Ptr{$T}, Ptr{$T}, Int64, Int64),
                handle().ptr, descriptor(Q.factors), Q.factors.data.ptr, Q.τ.ptr, B.ptr, ld(B), size(B, 2)))
    B
end
end # @eval
end # for T

# ---- structured +, -, alpha*A + B and A +- (I | Diagonal | Bidiagonal | Tridiagonal) on the device -------------------
# The result STRUCTURE stays with the reference (`similar(bc)`: src/broadcast.jl:107-131, 391-411 -- `similar` of a
# B200-backed matrix is B200-backed; copy_accommodating_diagonals: src/BlockSkylineMatrix.jl:502-531); these methods
# replace the per-block copyto! loops of src/broadcast.jl:317-388 and the element loops of src/BlockSkylineMatrix.jl:535-638.
skyptr(A::B200Sky) = A.data.ptr
bbbptr(A::B200BBB) = parent(A.data).ptr
for (T, suf) in ((Float64, "f64"), (Float32, "f32"))
    @eval function b200_axpby!(C::B200BBB{$T}, α::$T, A::B200BBB{$T}, β::$T, B::Union{Nothing,B200BBB{$T}})
        check(ccall(($(QuoteNode(Symbol("bbm_bbb_add_", suf))), libbbm), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, $T, Ptr{$T}, $T, Ptr{$T}, Ptr{$T}),
                    handle().ptr, descriptor(A), B === nothing ? C_NULL : descriptor(B), descriptor(C), α, bbbptr(A), β,
                    B === nothing ? Ptr{$T}(C_NULL) : bbbptr(B), bbbptr(C)))
        C
    end
    @eval function b200_axpby!(C::B200Sky{$T}, α::$T, A::Union{B200Sky{$T},B200BBB{$T}}, β::$T, B::Union{Nothing,B200Sky{$T},B200BBB{$T}})
        sky(X) = X isa B200Sky ? descriptor(X) : C_NULL
        bbb(X) = X isa B200BBB ? descriptor(X) : C_NULL
        ptr(X) = X === nothing ? Ptr{$T}(C_NULL) : (X isa B200Sky ? skyptr(X) : bbbptr(X))
        check(ccall(($(QuoteNode(Symbol("bbm_sky_add_", suf))), libbbm), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, $T, Ptr{$T}, $T, Ptr{$T}, Ptr{$T}),
                    handle().ptr, sky(A), bbb(A), sky(B), bbb(B), descriptor(C), α, ptr(A), β, ptr(B), skyptr(C)))
        C
    end
    # C .= A .+ B / C .= A .- B (src/broadcast.jl:317-372) and C .= α .* A .+ B (:379-388)
    for (op, β) in ((:+, 1), (:-, -1))
        @eval function Base.copyto!(C::Union{B200Sky{$T},B200BBB{$T}}, bc::Base.Broadcast.Broadcasted{<:BlockBandedMatrices.AbstractBlockBandedStyle,<:Any,typeof($op),
                                    <:Tuple{<:Union{B200Sky{$T},B200BBB{$T}},<:Union{B200Sky{$T},B200BBB{$T}}}})
            A, B = bc.args
            b200_axpby!(C, one($T), A, $T($β), B)
        end
    end
    @eval function Base.copyto!(C::Union{B200Sky{$T},B200BBB{$T}}, bc::Base.Broadcast.Broadcasted{<:BlockBandedMatrices.AbstractBlockBandedStyle,<:Any,typeof(+),
                                <:Tuple{<:Base.Broadcast.Broadcasted{<:BlockBandedMatrices.AbstractBlockBandedStyle,<:Any,typeof(*),<:Tuple{<:Number,<:Union{B200Sky{$T},B200BBB{$T}}}},
                                        <:Union{B200Sky{$T},B200BBB{$T}}}})
        αA, B = bc.args
        α, A = αA.args
        b200_axpby!(C, $T(α), A, one($T), B)
    end
    # in place: A[i,i] += α(λ + d[i]), A[i,i+1] += α du[i], A[i+1,i] += α dl[i]
    @eval function b200_add_diags!(A::B200Sky{$T}, α::$T, λ::$T, dl, d, du)
        dev(v) = v === nothing ? Ptr{$T}(C_NULL) : v.ptr
        check(ccall(($(QuoteNode(Symbol("bbm_sky_add_diags_", suf))), libbbm), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{$T}, $T, $T, Ptr{$T}, Ptr{$T}, Ptr{$T}),
                    handle().ptr, descriptor(A), skyptr(A), α, λ, dev(dl), dev(d), dev(du)))
        A
    end
    @eval function b200_add_diags!(A::B200BBB{$T}, α::$T, λ::$T, dl, d, du)
        dev(v) = v === nothing ? Ptr{$T}(C_NULL) : v.ptr
        check(ccall(($(QuoteNode(Symbol("bbm_bbb_add_diags_", suf))), libbbm), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{$T}, $T, $T, Ptr{$T}, Ptr{$T}, Ptr{$T}),
                    handle().ptr, descriptor(A), bbbptr(A), α, λ, dev(dl), dev(d), dev(du)))
        A
    end
    # A ± λI and λI ± A (src/BlockSkylineMatrix.jl:535-549); the Diagonal / Bidiagonal / Tridiagonal methods (:551-638)
    # follow the same two calls with their diagonals uploaded as B200Vectors
    for (op, s) in ((:+, 1), (:-, -1))
        @eval function Base.$op(A::B200Sky{$T}, J::UniformScaling)
            B = BlockBandedMatrices.copy_accommodating_diagonals(A, 0:0)      # `similar` keeps the B200 buffer type; the copy is b200_axpby!
            b200_add_diags!(B, $T($s), $T(J.λ), nothing, nothing, nothing)
        end
        @eval function Base.$op(J::UniformScaling, A::B200Sky{$T})
            B = BlockBandedMatrices.copy_accommodating_diagonals(A, 0:0)
            $s == -1 && LinearAlgebra.lmul!(-one($T), B.data)
            b200_add_diags!(B, one($T), $T(J.λ), nothing, nothing, nothing)
        end
    end
    @eval function BlockBandedMatrices.BlockSkylineMatrix{$T}(A::B200Sky{$T}, sizes::BlockBandedMatrices.BlockSkylineSizes)
        B = BlockBandedMatrices._BlockSkylineMatrix(B200Array{$T}(undef, BlockBandedMatrices.bb_numentries(sizes)), sizes)
        b200_axpby!(B, one($T), A, zero($T), nothing)                        # copy into the wider structure (:530)
    end
    @eval function LinearAlgebra.lmul!(a::$T, x::B200Array{$T})
        check(ccall(($(QuoteNode(Symbol("bbm_scal_", suf))), libbbm), Int32, (Ptr{Cvoid}, Int64, $T, Ptr{$T}), handle().ptr, length(x), a, x.ptr))
        x
    end
end

# ---- several GPUs (one Julia process per GPU, MPI.jl / Distributed ships the 128-byte NCCL id) --------------------------
# bbm_comm_unique_id / bbm_comm_create, then per matrix structure one plan:
#   banded-block-banded matvec        bbm_dist_bbb_plan_create / bbm_dist_bbb_enable_peer / bbm_dist_bbb_mul_f64
#   block-banded matvec               bbm_dist_sky_plan_create / bbm_dist_sky_mul_f64
#   block-banded x block-banded       bbm_dist_sky_mm_plan_create / bbm_dist_sky_matmul_f64   (cfg4: row blocks over 8 GPUs)
# The local structures are ordinary B200Sky / B200BBB matrices built from bbm_dist_sky_plan_layout /
# bbm_dist_sky_mm_plan_info + *_bandwidths (INTEGRATION.md shows the calls).
function dist_sky_mul!(plan::Ptr{Cvoid}, y::B200Array{Float64}, Aloc::B200Sky{Float64}, xlocal::B200Array{Float64}, α=1.0, β=0.0)
    check(ccall((:bbm_dist_sky_mul_f64, libbbm), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Int64, Float64, Ptr{Float64}),
                plan, α, skyptr(Aloc), xlocal.ptr, size(xlocal, 2), β, y.ptr))
    y
end
function dist_sky_matmul!(plan::Ptr{Cvoid}, Cloc::B200Sky{Float64}, Aloc::B200Sky{Float64}, Bloc::B200Sky{Float64}, α=1.0, β=0.0)
    check(ccall((:bbm_dist_sky_matmul_f64, libbbm), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}),
                plan, α, skyptr(Aloc), skyptr(Bloc), β, skyptr(Cloc)))
    Cloc
end

synchronize() = check(ccall((:bbm_synchronize, libbbm), Int32, (Ptr{Cvoid},), handle().ptr))

end # module
