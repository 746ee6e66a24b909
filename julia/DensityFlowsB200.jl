# DensityFlowsB200.jl -- thin `ccall` layer that lets DensityFlows.jl run its coupling-chain hot path on libdflow.so
# (hand-written sm_100a kernels, include/dflow.h).  Load AFTER `using DensityFlows, CUDA`.
#
# NOTE: Julia is not installed in the build environment of this repository, so this file has never been executed;
# every call below mirrors, argument for argument, the ctypes binding in densityflows.jl_b200/_lib.py + model.py,
# which IS exercised by tests/ on a B200.  Treat it as the reference-side stub a maintainer adapts (INTEGRATION.md).
#
# What it adds (no method of the reference is overwritten for CPU arrays):
#   backward(chain, x::CuArray, θ::CuArray)        -> dflow_normalize          (src/Chains.jl:149-164)
#   forward(chain, z::CuArray, θ::CuArray)         -> dflow_forward_ldj        (src/Chains.jl:167-183)
#   forward!(chain, z::CuArray, θ::CuArray)        -> dflow_sample_inplace     (src/Chains.jl:187-197)
#   logpdf(flow, x::CuArray, θ::CuArray)           -> dflow_logpdf             (src/Flows.jl:272-281)
#   sample(flow, dims, θ::NTuple; device=true)     -> dflow_sample_rng         (src/Flows.jl:157-192)
#   train!(flow, data::CuDataArrays, state; ...)   -> dflow_loss_grad + dflow_adam_step (src/Flows.jl:380-445)
#   ChainRulesCore.rrule(backward, chain, x, θ)    -> dflow_loss_grad-style pullback for Zygote users
module DensityFlowsB200

using DensityFlows
using CUDA
import DensityFlows: backward, forward, forward!, logpdf, sample, train!
import DensityFlows: FlowChain, CouplingBlock, RNVPCouplingLayer, NICECouplingLayer, NormalizationLayer, Flow, DataArrays
import Flux

const libdflow = get(ENV, "LIBDFLOW", joinpath(@__DIR__, "..", "densityflows.jl_b200", "lib", "libdflow.so"))

# ---- C structs (include/dflow.h) ---------------------------------------------------------------------------
struct NetDesc
    depth::Int32
    widths::Ptr{Int32}
    acts::Ptr{Int32}
    has_bias::Int32
end

struct ElemDesc
    kind::Int32
    n_af::Int32
    axis_af::Ptr{Int32}
    n_id::Int32
    axis_id::Ptr{Int32}
    s_net::NetDesc
    t_net::NetDesc
    x_min::Ptr{Float32}
    x_max::Ptr{Float32}
    alpha::Float32
    beta::Float32
end

struct ChainDesc
    d::Int32
    n::Int32
    n_elems::Int32
    elems::Ptr{ElemDesc}
    theta_min::Ptr{Float32}
    theta_max::Ptr{Float32}
end

const ELEM_RNVP, ELEM_NICE, ELEM_NORM = Int32(0), Int32(1), Int32(2)
const THETA_NORMALIZE = Int32(1)

act_code(f) = f === Flux.relu ? Int32(1) : f === tanh ? Int32(2) : f === Flux.sigmoid ? Int32(3) :
              f === identity ? Int32(0) : throw(ArgumentError("activation $f has no fused kernel"))

function check(rc::Integer)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:dflow_last_error, libdflow), Cstring, ()))
    # DFLOW_E_INVALID_ARG / DFLOW_E_UNSUPPORTED map to the reference's ArgumentError (src/Blocks.jl:71, src/Flows.jl:222)
    (rc == -1 || rc == -2) ? throw(ArgumentError(msg)) : error("libdflow error $rc: $msg")
end

# ---- flatten a FlowChain into leaf elements in chain order (blocks -> layer_1, layer_2; src/Blocks.jl:127-161) ----
leaves(e::RNVPCouplingLayer) = Any[e]
leaves(e::NICECouplingLayer) = Any[e]
leaves(e::NormalizationLayer) = Any[e]
leaves(b::CouplingBlock) = Any[b.layer_1, b.layer_2]
leaves(c::FlowChain) = reduce(vcat, (leaves(l) for l in c.layers); init = Any[])

nets(e::RNVPCouplingLayer) = (e.s_net, e.t_net)      # Flux.@layer ... trainable=(s_net, t_net), src/affine/RNVP.jl:51
nets(e::NICECouplingLayer) = (e.t_net,)
nets(::NormalizationLayer) = ()

# ---- the packed device twin of a chain --------------------------------------------------------------------
mutable struct PackedChain
    handle::Ptr{Cvoid}
    d::Int
    n::Int
    P::Int
    W::CuVector{Float32}            # packed parameters: chain order, s_net then t_net, vec(weight) then bias
    ws::CuVector{UInt8}             # adjoint workspace (dflow_workspace_bytes)
    leaves::Vector{Any}
end

function net_desc(net::Flux.Chain, keep)
    widths = Int32[size(net.layers[1].weight, 2); [size(l.weight, 1) for l in net.layers]...]
    acts = Int32[act_code(l.σ) for l in net.layers]
    push!(keep, widths, acts)
    NetDesc(Int32(length(net.layers)), pointer(widths), pointer(acts), Int32(net.layers[1].bias !== false))
end

const NULLNET = NetDesc(0, C_NULL, C_NULL, 0)

function PackedChain(chain::FlowChain; θ_min = nothing, θ_max = nothing)
    ls = leaves(chain)
    keep = Any[]
    d = n = 0
    elems = ElemDesc[]
    for e in ls
        if e isa NormalizationLayer
            xmin, xmax = Float32.(collect(e.x_min)), Float32.(collect(e.x_max))
            push!(keep, xmin, xmax)
            d = length(xmin)
            push!(elems, ElemDesc(ELEM_NORM, 0, C_NULL, 0, C_NULL, NULLNET, NULLNET, pointer(xmin), pointer(xmax),
                                  Float32(e.α), Float32(e.β)))
        else
            d, n = e.axes.d, e.axes.n
            af = Int32.(e.axes.axis_af .- 1)          # 1-based -> 0-based at the C boundary, caller order kept
            id = Int32.(e.axes.axis_id .- 1)          # explicit: reverse(axes) keeps an un-sorted axis_id
            push!(keep, af, id)
            s = e isa RNVPCouplingLayer ? net_desc(e.s_net, keep) : NULLNET
            t = net_desc(e.t_net, keep)
            push!(elems, ElemDesc(e isa RNVPCouplingLayer ? ELEM_RNVP : ELEM_NICE, Int32(length(af)), pointer(af),
                                  Int32(length(id)), pointer(id), s, t, C_NULL, C_NULL, 0f0, 0f0))
        end
    end
    tmin = θ_min === nothing ? Float32[] : Float32.(collect(θ_min))
    tmax = θ_max === nothing ? Float32[] : Float32.(collect(θ_max))
    desc = Ref(ChainDesc(Int32(d), Int32(n), Int32(length(elems)), pointer(elems),
                         isempty(tmin) ? C_NULL : pointer(tmin), isempty(tmax) ? C_NULL : pointer(tmax)))
    handle = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep elems tmin tmax begin
        check(ccall((:dflow_chain_create, libdflow), Cint, (Ref{ChainDesc}, Ref{Ptr{Cvoid}}), desc, handle))
    end
    P = Int(ccall((:dflow_param_count, libdflow), Int64, (Ptr{Cvoid},), handle[]))
    pc = PackedChain(handle[], d, n, P, CUDA.zeros(Float32, max(P, 1)), CuVector{UInt8}(undef, 0), ls)
    finalizer(p -> ccall((:dflow_chain_destroy, libdflow), Cint, (Ptr{Cvoid},), p.handle), pc)
    pack!(pc)
    return pc
end

"Copy the Flux parameters into the packed device buffer (weights stay authoritative in the Flux structs)."
function pack!(pc::PackedChain)
    host = Float32[]
    for e in pc.leaves, net in nets(e), l in net.layers
        append!(host, vec(l.weight))                 # Flux (out,in) column-major, verbatim
        l.bias === false || append!(host, l.bias)
    end
    @assert length(host) == pc.P
    copyto!(pc.W, 1, host, 1, pc.P)
    return pc
end

"Copy the packed device buffer back into the Flux structs (after train!)."
function unpack!(pc::PackedChain)
    host = Array(pc.W)
    off = 0
    for e in pc.leaves, net in nets(e), l in net.layers
        k = length(l.weight); copyto!(l.weight, reshape(view(host, off+1:off+k), size(l.weight))); off += k
        if l.bias !== false
            k = length(l.bias); copyto!(l.bias, view(host, off+1:off+k)); off += k
        end
    end
    return pc
end

const _cache = IdDict{Any,PackedChain}()
packed(chain::FlowChain) = get!(() -> PackedChain(chain), _cache, chain)
packed(flow::Flow) = get!(() -> PackedChain(flow.model; θ_min = flow.metadata.θ_min, θ_max = flow.metadata.θ_max), _cache, flow)

nsamples(x) = prod(size(x)[2:end])
stream() = Base.unsafe_convert(Ptr{Cvoid}, CUDA.stream().handle)
θptr(θ::CuArray) = size(θ, 1) == 0 ? CU_NULL : pointer(θ)

# ---- element protocol on device arrays (src/Chains.jl:33-72) --------------------------------------------------
function _normalize(pc, x::CuArray{Float32}, θ::CuArray{Float32}, flags)
    z = similar(x); ldj = CuArray{Float32}(undef, size(x)[2:end]...)
    check(ccall((:dflow_normalize, libdflow), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Int32, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
                pc.handle, pc.W, x, θptr(θ), nsamples(x), flags, z, ldj, stream()))
    return z, ldj
end
backward(chain::FlowChain, x::CuArray{Float32,N}, θ::CuArray{Float32,N}) where {N} = _normalize(packed(chain), x, θ, Int32(0))
backward(flow::Flow{Float32}, x::CuArray{Float32,N}, θ::CuArray{Float32,N}) where {N} =
    _normalize(packed(flow), x, θ, flow.metadata.n > 0 ? THETA_NORMALIZE : Int32(0))   # @flow_wrapper, src/Macros.jl:104-112

function _forward(pc, z::CuArray{Float32}, θ::CuArray{Float32}, flags)
    x = similar(z); ldj = CuArray{Float32}(undef, size(z)[2:end]...)
    check(ccall((:dflow_forward_ldj, libdflow), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Int32, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
                pc.handle, pc.W, z, θptr(θ), nsamples(z), flags, x, ldj, stream()))
    return x, ldj
end
forward(chain::FlowChain, z::CuArray{Float32,N}, θ::CuArray{Float32,N}) where {N} = _forward(packed(chain), z, θ, Int32(0))
forward(flow::Flow{Float32}, z::CuArray{Float32,N}, θ::CuArray{Float32,N}) where {N} =
    _forward(packed(flow), z, θ, flow.metadata.n > 0 ? THETA_NORMALIZE : Int32(0))

function _forward!(pc, z::CuArray{Float32}, θ::CuArray{Float32}, flags)
    check(ccall((:dflow_sample_inplace, libdflow), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Int32, Ptr{Cvoid}),
                pc.handle, pc.W, z, θptr(θ), CU_NULL, nsamples(z), flags, stream()))
    return nothing
end
forward!(chain::FlowChain, z::CuArray{Float32}, θ::CuArray{Float32}) = _forward!(packed(chain), z, θ, Int32(0))
forward!(flow::Flow{Float32}, z::CuArray{Float32}, θ::CuArray{Float32}) =
    _forward!(packed(flow), z, θ, flow.metadata.n > 0 ? THETA_NORMALIZE : Int32(0))

# ---- logpdf / sample (src/Flows.jl:157-192, 272-281) ----------------------------------------------------------
function logpdf(flow::Flow{Float32}, x::CuArray{Float32,N}, θ::CuArray{Float32,N}) where {N}
    pc = packed(flow); out = CuArray{Float32}(undef, size(x)[2:end]...)
    check(ccall((:dflow_logpdf, libdflow), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, CuPtr{Int32}, Int32, CuPtr{Float32}, Ptr{Cvoid}),
                pc.handle, pc.W, x, θptr(θ), nsamples(x), CU_NULL, flow.metadata.n > 0 ? THETA_NORMALIZE : Int32(0), out, stream()))
    return out
end

"sample(flow, dims, θ::NTuple; seed) on the device: base draw (Philox4x32-10) + forward! fused in one kernel."
function sample_device(flow::Flow{Float32,D}, dims::Tuple{Vararg{Integer}}, θ::NTuple{Nθ,Float32};
                       seed::UInt64 = rand(UInt64)) where {D,Nθ}
    pc = packed(flow); B = prod(dims); out = CuArray{Float32}(undef, D, dims...)
    θc = CuArray(collect(θ))
    check(ccall((:dflow_sample_rng, libdflow), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, UInt64, UInt32, UInt64, CuPtr{Float32}, CuPtr{Float32}, Int64, Int32, CuPtr{Float32}, Ptr{Cvoid}),
                pc.handle, pc.W, seed, UInt32(0), UInt64(0), CU_NULL, Nθ > 0 ? pointer(θc) : CU_NULL, B,
                Nθ > 0 ? THETA_NORMALIZE : Int32(0), out, stream()))
    return out
end

"""
sample_with_rejection(condition, flow, dims, θ::NTuple, m = 100) on the device (src/Flows.jl:196-229): the same stream
of points as `sample_device` with that seed (Philox counter = draw index), drawn in batches; `condition(points, θ)`
gets a `(D, nb)` CuArray and returns a Bool vector over the batch; accepted points are compacted in draw order.
"""
function sample_with_rejection_device(condition::Function, flow::Flow{Float32,D}, dims::Tuple{Vararg{Integer}},
                                      θ::NTuple{Nθ,Float32}, m::Int = 100; seed::UInt64 = rand(UInt64)) where {D,Nθ}
    pc = packed(flow); n = prod(dims); out = CuArray{Float32}(undef, D, n)
    θc = CuArray(collect(θ))
    have = 0; drawn = 0; nb = max(1024, 2n)
    while have < n && drawn < m * n
        cur = min(nb, m * n - drawn)
        pts = CuArray{Float32}(undef, D, cur)
        check(ccall((:dflow_sample_rng, libdflow), Cint,
                    (Ptr{Cvoid}, CuPtr{Float32}, UInt64, UInt32, UInt64, CuPtr{Float32}, CuPtr{Float32}, Int64, Int32, CuPtr{Float32}, Ptr{Cvoid}),
                    pc.handle, pc.W, seed, UInt32(0), UInt64(drawn), CU_NULL, Nθ > 0 ? pointer(θc) : CU_NULL, cur,
                    Nθ > 0 ? THETA_NORMALIZE : Int32(0), pts, stream()))
        keep = pts[:, findall(Array(condition(pts, θ)))]
        take = min(size(keep, 2), n - have)
        out[:, have+1:have+take] .= keep[:, 1:take]
        have += take; drawn += cur; nb = min(2nb, 1 << 24)
    end
    have < n && throw(ArgumentError("Impossible to reach convergence of rejection sampling"))
    return reshape(out, D, dims...)
end

# ---- train! on device-resident data (src/Flows.jl:380-445) -----------------------------------------------------
"Device-resident DataArrays: x, θ on the GPU and the partition as 0-based Int32 index vectors."
struct CuDataArrays
    x::CuMatrix{Float32}
    θ::CuMatrix{Float32}
    training::CuVector{Int32}
    validation::CuVector{Int32}
end
CuDataArrays(data::DataArrays) = CuDataArrays(CuArray(data.x), CuArray(data.θ),
                                              CuArray(Int32.(data.partition.training .- 1)),
                                              CuArray(Int32.(data.partition.validation .- 1)))

mutable struct AdamState
    m::CuVector{Float32}; v::CuVector{Float32}; t::Int; η::Float32; β::NTuple{2,Float32}; ϵ::Float32
end
AdamState(pc::PackedChain; η = 1f-3, β = (0.9f0, 0.999f0), ϵ = 1f-8) =
    AdamState(CUDA.zeros(Float32, max(pc.P, 1)), CUDA.zeros(Float32, max(pc.P, 1)), 0, η, β, ϵ)

function _full_loss(pc, data, idx, flags)
    acc = CUDA.zeros(Float32, 2)
    check(ccall((:dflow_logpdf_sum, libdflow), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, CuPtr{Int32}, Int32, CuPtr{Float32}, Ptr{Cvoid}),
                pc.handle, pc.W, data.x, θptr(data.θ), length(idx), idx, flags, acc, stream()))
    s, bad = Array(acc)
    return bad > 0 ? NaN32 : -s / length(idx)       # loss = -mean(logpdf + ldj), src/Flows.jl:352-359
end

function train!(flow::Flow{Float32}, data::CuDataArrays, st::AdamState; epochs::Int = 100, batchsize::Int = 64,
                shuffle::Bool = true, verbose::Bool = true)
    pc = packed(flow); flags = flow.metadata.n > 0 ? THETA_NORMALIZE : Int32(0)
    need = ccall((:dflow_workspace_bytes, libdflow), Csize_t, (Ptr{Cvoid}, Int64), pc.handle, batchsize)
    length(pc.ws) < need && (pc.ws = CuVector{UInt8}(undef, need))
    buf = CUDA.zeros(Float32, max(pc.P, 1) + 2); grad = view(buf, 1:max(pc.P, 1)); loss2 = view(buf, max(pc.P, 1)+1:max(pc.P, 1)+2)
    ntr = length(data.training)
    for _ in 1:epochs
        order = shuffle ? data.training[CuArray(Int32.(Random.randperm(ntr)))] : data.training
        for b0 in 1:batchsize:ntr                    # Flux.DataLoader: partial last batch kept
            idx = view(order, b0:min(b0 + batchsize - 1, ntr)); nb = length(idx)
            fill!(buf, 0f0)
            check(ccall((:dflow_loss_grad, libdflow), Cint,
                        (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, CuPtr{Int32}, Float32, Int32,
                         CuPtr{Float32}, CuPtr{Float32}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                        pc.handle, pc.W, data.x, θptr(data.θ), nb, idx, 1f0 / nb, flags, loss2, grad, pc.ws, length(pc.ws), stream()))
            st.t += 1
            check(ccall((:dflow_adam_step, libdflow), Cint,
                        (CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Float32, Float32, Float32, Float32, Int64, Ptr{Cvoid}),
                        pc.W, grad, st.m, st.v, pc.P, st.η, st.β[1], st.β[2], st.ϵ, st.t, stream()))
        end
        push!(flow.train_loss, _full_loss(pc, data, data.training, flags))       # src/Flows.jl:419-421
        push!(flow.valid_loss, _full_loss(pc, data, data.validation, flags))     # src/Flows.jl:428-430
        verbose && println("epoch: $(length(flow.train_loss)) | train_loss = $(flow.train_loss[end]), valid_loss = $(flow.valid_loss[end])")
    end
    unpack!(pc)                                       # Flux structs see the trained weights
    return nothing
end

# ---- data-parallel step over NVLink peer memory (include/dflow.h: dflow_dp_*) -------------------------------------
# One Julia process per GPU (e.g. MPI.jl).  `exchange(handle::Vector{UInt8})::Vector{UInt8}` must return the ranks'
# 64-byte IPC handles concatenated in rank order (an MPI.Allgather of 64 bytes).  Every step accumulates the local
# gradient straight into the communication buffer and ONE kernel per rank reduces the peers' buffers and applies Adam.
mutable struct PeerStep
    dp::Ptr{Cvoid}
    P::Int
end

function PeerStep(pc, rank::Integer, nranks::Integer, exchange)
    dp = Ref{Ptr{Cvoid}}(C_NULL); handle = zeros(UInt8, 64)
    check(ccall((:dflow_dp_create, libdflow), Cint, (Int32, Int32, Int64, Ptr{Ptr{Cvoid}}, Ptr{UInt8}),
                rank, nranks, max(pc.P, 1), dp, handle))
    check(ccall((:dflow_dp_connect, libdflow), Cint, (Ptr{Cvoid}, Ptr{UInt8}), dp[], exchange(handle)))
    return PeerStep(dp[], max(pc.P, 1))
end

function peer_train_step!(ps::PeerStep, pc, st::AdamState, x, θ, idx, B_global::Integer, flags::Int32, loss2)
    buf = ccall((:dflow_dp_grad_buffer, libdflow), CuPtr{Float32}, (Ptr{Cvoid},), ps.dp)
    CUDA.memset(buf, UInt32(0), ps.P + 2)              # [grad | Σlogp | #non-finite] of this step
    nb = length(idx)
    need = ccall((:dflow_workspace_bytes, libdflow), Csize_t, (Ptr{Cvoid}, Int64), pc.handle, nb)
    length(pc.ws) < need && (pc.ws = CuVector{UInt8}(undef, need))
    check(ccall((:dflow_loss_grad, libdflow), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, CuPtr{Int32}, Float32, Int32,
                 CuPtr{Float32}, CuPtr{Float32}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                pc.handle, pc.W, x, θptr(θ), nb, idx, 1f0 / B_global, flags, buf + 4 * ps.P, buf, pc.ws, length(pc.ws), stream()))
    st.t += 1
    check(ccall((:dflow_dp_allreduce_adam, libdflow), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Float32, Float32, Float32, Float32, Int64,
                 CuPtr{Float32}, Ptr{Cvoid}),
                ps.dp, pc.W, st.m, st.v, st.η, st.β[1], st.β[2], st.ϵ, st.t, loss2, stream()))
    return nothing
end

# tuning knobs (dflow_set_tuning): "tc_mode" 1 / 0 / -1 forces an eligible hidden <= 64 chain onto / off the tensor cores,
# "tc_ws_budget_mb" caps the adjoint workspace, "tc_cluster" 0 / 1 / 2 selects independent CTAs / multicast pairs /
# cta_group::2 pairs for streamed conditioners.
tune!(pc, key::AbstractString, value::Integer) =
    check(ccall((:dflow_set_tuning, libdflow), Cint, (Ptr{Cvoid}, Cstring, Int32), pc.handle, key, value))

end # module
