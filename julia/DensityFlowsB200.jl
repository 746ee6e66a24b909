# DensityFlowsB200.jl -- thin `ccall` layer that lets DensityFlows.jl run its coupling-chain hot path on libdflow.so
# (hand-written sm_100a kernels, include/dflow.h).  Load AFTER `using DensityFlows, CUDA`.
#
# STATUS: Julia is not installed in the build environment of this repository (no `julia` binary, no network), so this file
# has NEVER been executed.  Every ccall below mirrors, argument for argument, the ctypes binding in
# densityflows.jl_b200/_lib.py + model.py + flows.py, which IS exercised by tests/ on a B200 through the same C ABI.
# Treat it as the reference-side stub a maintainer adapts (INTEGRATION.md).
#
# It only ADDS methods, all on the reference's own types, selected by dispatch on device arrays (no CPU method of the
# reference is overwritten):
#   backward / forward / forward!(::FlowChain | ::Flow, ::CuArray, ::CuArray)   src/Chains.jl:149-197, src/Macros.jl:104-112
#   logpdf(::Flow, x::CuArray, θ::CuArray | ::NTuple)                           src/Flows.jl:272-284
#   logpdf(DeviceGrid(), ::Flow, x::NTuple{D,Vector}, θ::NTuple)                src/Flows.jl:287-331 (in-kernel grid)
#   sample(rng::DeviceRNG, ::Flow, dims, θ::NTuple | ::CuArray)                 src/Flows.jl:157-192
#   train!(::Flow, ::DataArrays{Float32,2,<:CuArray,...}, state; kws...)         src/Flows.jl:380-445
#   ChainRulesCore.rrule(::typeof(backward), ::FlowChain, ::CuArray, ::CuArray)  src/affine/RNVP.jl:99-147 through the chain
#   train_local!(flow, shards, state; devices, kws...)                          one process, several GPUs (dflow_dp_*_local)
module DensityFlowsB200

using DensityFlows
using CUDA
using Random
import ChainRulesCore
import ChainRulesCore: Tangent, NoTangent
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

# dflow_dp_shard: one rank's share of a minibatch for dflow_dp_train_step
struct DpShard
    chain::Ptr{Cvoid}
    W::CuPtr{Float32}
    m::CuPtr{Float32}
    v::CuPtr{Float32}
    x::CuPtr{Float32}
    theta::CuPtr{Float32}
    B::Int64
    idx::CuPtr{Int32}
    ws::CuPtr{Cvoid}
    ws_bytes::Csize_t
    loss2_out::CuPtr{Float32}
    stream::Ptr{Cvoid}
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
    ws::CuVector{UInt8}             # adjoint workspace (dflow_workspace_bytes), caller owned
    scratch::CuVector{UInt8}        # forward-call scratch (dflow_scratch_bytes), caller owned
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
    pc = PackedChain(handle[], d, n, P, CUDA.zeros(Float32, max(P, 1)), CuVector{UInt8}(undef, 0), CuVector{UInt8}(undef, 0), ls)
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
flagsof(flow::Flow) = flow.metadata.n > 0 ? THETA_NORMALIZE : Int32(0)   # @flow_wrapper: θ normalised with the flow's range

"Caller-owned scratch of the forward-type calls (0 bytes on the CUDA-core kernels); the library never allocates."
function ensure_scratch!(pc::PackedChain, B::Integer)
    need = ccall((:dflow_scratch_bytes, libdflow), Csize_t, (Ptr{Cvoid}, Int64), pc.handle, B)
    if need > length(pc.scratch)
        CUDA.synchronize()
        pc.scratch = CuVector{UInt8}(undef, need)
        check(ccall((:dflow_chain_set_scratch, libdflow), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, Csize_t), pc.handle, pc.scratch, need))
    end
end
function workspace!(pc::PackedChain, B::Integer)
    need = ccall((:dflow_workspace_bytes, libdflow), Csize_t, (Ptr{Cvoid}, Int64), pc.handle, B)
    length(pc.ws) < need && (pc.ws = CuVector{UInt8}(undef, need))
    return pc.ws
end

# ---- element protocol on device arrays (src/Chains.jl:33-72) --------------------------------------------------
function _normalize(pc, x::CuArray{Float32}, θ::CuArray{Float32}, flags)
    z = similar(x); ldj = CuArray{Float32}(undef, size(x)[2:end]...)
    ensure_scratch!(pc, nsamples(x))
    check(ccall((:dflow_normalize, libdflow), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Int32, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
                pc.handle, pc.W, x, θptr(θ), nsamples(x), flags, z, ldj, stream()))
    return z, ldj
end
backward(chain::FlowChain, x::CuArray{Float32,N}, θ::CuArray{Float32,N}) where {N} = _normalize(packed(chain), x, θ, Int32(0))
backward(flow::Flow{Float32}, x::CuArray{Float32,N}, θ::CuArray{Float32,N}) where {N} = _normalize(packed(flow), x, θ, flagsof(flow))

function _forward(pc, z::CuArray{Float32}, θ::CuArray{Float32}, flags)
    x = similar(z); ldj = CuArray{Float32}(undef, size(z)[2:end]...)
    ensure_scratch!(pc, nsamples(z))
    check(ccall((:dflow_forward_ldj, libdflow), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Int32, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
                pc.handle, pc.W, z, θptr(θ), nsamples(z), flags, x, ldj, stream()))
    return x, ldj
end
forward(chain::FlowChain, z::CuArray{Float32,N}, θ::CuArray{Float32,N}) where {N} = _forward(packed(chain), z, θ, Int32(0))
forward(flow::Flow{Float32}, z::CuArray{Float32,N}, θ::CuArray{Float32,N}) where {N} = _forward(packed(flow), z, θ, flagsof(flow))

function _forward!(pc, z::CuArray{Float32}, θ::CuArray{Float32}, flags)
    ensure_scratch!(pc, nsamples(z))
    check(ccall((:dflow_sample_inplace, libdflow), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Int32, Ptr{Cvoid}),
                pc.handle, pc.W, z, θptr(θ), CU_NULL, nsamples(z), flags, stream()))
    return nothing
end
forward!(chain::FlowChain, z::CuArray{Float32}, θ::CuArray{Float32}) = _forward!(packed(chain), z, θ, Int32(0))
forward!(flow::Flow{Float32}, z::CuArray{Float32}, θ::CuArray{Float32}) = _forward!(packed(flow), z, θ, flagsof(flow))

# ---- logpdf (src/Flows.jl:272-331) ------------------------------------------------------------------------------
function logpdf(flow::Flow{Float32}, x::CuArray{Float32,N}, θ::CuArray{Float32,N}) where {N}
    pc = packed(flow); out = CuArray{Float32}(undef, size(x)[2:end]...)
    ensure_scratch!(pc, nsamples(x))
    check(ccall((:dflow_logpdf, libdflow), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, CuPtr{Int32}, Int32, CuPtr{Float32}, Ptr{Cvoid}),
                pc.handle, pc.W, x, θptr(θ), nsamples(x), CU_NULL, flagsof(flow), out, stream()))
    return out
end
# fixed condition for every column: the reference collects θ and broadcasts it (src/Flows.jl:284)
logpdf(flow::Flow{Float32,D,Nθ}, x::CuArray{Float32}, θ::NTuple{Nθ,Float32}) where {D,Nθ} =
    logpdf(flow, x, CuArray(repeat(collect(θ), 1, size(x)[2:end]...)))

"Dispatch tag: evaluate the grid methods of logpdf on the device."
struct DeviceGrid end

"""
    logpdf(DeviceGrid(), flow, x::NTuple{D,Vector}, θ::NTuple)

src/Flows.jl:287-331 on the device: values on the tensor-product grid of the d coordinate vectors (first vector fastest,
like Iterators.product); the (d, prod(lens)) point array and the broadcast θ are never materialised (dflow_logpdf_grid).
"""
function logpdf(::DeviceGrid, flow::Flow{Float32,D,Nθ}, x::NTuple{D,AbstractVector{Float32}}, θ::NTuple{Nθ,Float32} = ()) where {D,Nθ}
    pc = packed(flow)
    lens = Int64[length(v) for v in x]
    vals = CuArray(reduce(vcat, collect.(x)))
    θc = CuArray(Float32[θ...])
    out = CuArray{Float32}(undef, lens...)
    ensure_scratch!(pc, prod(lens))
    check(ccall((:dflow_logpdf_grid, libdflow), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, Ptr{Int64}, CuPtr{Float32}, Int32, CuPtr{Float32}, Ptr{Cvoid}),
                pc.handle, pc.W, vals, lens, Nθ > 0 ? pointer(θc) : CU_NULL, flagsof(flow), out, stream()))
    return out
end

# ---- sample (src/Flows.jl:157-192) ------------------------------------------------------------------------------
"""
    DeviceRNG(seed)

RNG tag for `sample(rng, flow, dims, θ)`: the base draw r ~ N(0, I) happens inside the sampling kernel
(Philox4x32-10 + Box-Muller, counter = sample index: dflow_sample_rng) instead of `rand(rng, flow.base, n)` + `forward!`.
Every call advances `offset`, so successive calls draw fresh points.
"""
mutable struct DeviceRNG <: Random.AbstractRNG
    seed::UInt64
    offset::UInt32
end
DeviceRNG(seed::Integer = rand(UInt64)) = DeviceRNG(UInt64(seed), UInt32(0))

function _sample_rng(flow::Flow{Float32,D}, rng::DeviceRNG, dims, θ, θc; first::Integer = 0) where {D}
    pc = packed(flow); B = prod(dims); out = CuArray{Float32}(undef, D, dims...)
    ensure_scratch!(pc, B)
    check(ccall((:dflow_sample_rng, libdflow), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, UInt64, UInt32, UInt64, CuPtr{Float32}, CuPtr{Float32}, Int64, Int32, CuPtr{Float32}, Ptr{Cvoid}),
                pc.handle, pc.W, rng.seed, rng.offset, UInt64(first), θ, θc, B, flagsof(flow), out, stream()))
    return out
end

function sample(rng::DeviceRNG, flow::Flow{Float32,D,Nθ}, dims::Tuple{Vararg{Integer}}, θ::NTuple{Nθ,Float32}) where {D,Nθ}
    θc = CuArray(Float32[θ...])
    out = _sample_rng(flow, rng, dims, CU_NULL, Nθ > 0 ? pointer(θc) : CU_NULL)
    rng.offset += UInt32(1)
    return out
end
function sample(rng::DeviceRNG, flow::Flow{Float32,D}, dims::NTuple{M,Integer}, θ::CuArray{Float32,K}) where {D,M,K}
    @assert K == M + 1 "dimensions θ must match (n, dims...) with n number of trained parameters"   # src/Flows.jl:165
    out = _sample_rng(flow, rng, dims, θptr(θ), CU_NULL)
    rng.offset += UInt32(1)
    return out
end
sample(rng::DeviceRNG, flow::Flow{Float32}, dims::Integer, θ) = sample(rng, flow, (dims,), θ)

"""
sample_with_rejection(rng::DeviceRNG, condition, flow, dims, θ::NTuple, m = 100) (src/Flows.jl:196-229): the same stream of
points as `sample(rng, ...)` (Philox counter = draw index), drawn in batches; `condition(points, θ)` gets a `(D, nb)`
CuArray and returns a Bool vector over the batch; accepted points are compacted in draw order.
"""
function DensityFlows.sample_with_rejection(rng::DeviceRNG, condition::Function, flow::Flow{Float32,D,Nθ},
                                            dims::Tuple{Vararg{Integer}}, θ::NTuple{Nθ,Float32}, m::Int = 100) where {D,Nθ}
    n = prod(dims); out = CuArray{Float32}(undef, D, n)
    θc = CuArray(Float32[θ...])
    have = 0; drawn = 0; nb = max(1024, 2n)
    while have < n && drawn < m * n
        cur = min(nb, m * n - drawn)
        pts = _sample_rng(flow, rng, (cur,), CU_NULL, Nθ > 0 ? pointer(θc) : CU_NULL; first = drawn)
        keep = pts[:, findall(Array(condition(pts, θ)))]
        take = min(size(keep, 2), n - have)
        out[:, have+1:have+take] .= keep[:, 1:take]
        have += take; drawn += cur; nb = min(2nb, 1 << 24)
    end
    rng.offset += UInt32(1)
    have < n && throw(ArgumentError("Impossible to reach convergence of rejection sampling"))   # src/Flows.jl:221-224
    return reshape(out, D, dims...)
end

# ---- the adjoint: rrule of backward(chain, x, θ) (src/affine/RNVP.jl:99-147 composed through src/Chains.jl:149-164) ----
"Parameter cotangent in the packed layout -> Tangent mirroring the Flux.@layer structure of the chain."
function tangent_of(chain::FlowChain, g::Vector{Float32})
    off = Ref(0)
    take(dims...) = (k = prod(dims); r = reshape(g[off[]+1:off[]+k], dims...); off[] += k; r)
    dense(l) = Tangent{typeof(l)}(weight = take(size(l.weight)...), bias = l.bias === false ? NoTangent() : take(length(l.bias)))
    net(c::Flux.Chain) = Tangent{typeof(c)}(layers = map(dense, c.layers))
    elem(e::RNVPCouplingLayer) = Tangent{typeof(e)}(s_net = net(e.s_net), t_net = net(e.t_net))
    elem(e::NICECouplingLayer) = Tangent{typeof(e)}(t_net = net(e.t_net))
    elem(::NormalizationLayer) = NoTangent()                      # not trainable (src/norm/Normalization.jl:61)
    elem(b::CouplingBlock) = (t1 = elem(b.layer_1); t2 = elem(b.layer_2); Tangent{typeof(b)}(layer_1 = t1, layer_2 = t2))
    elem(c::FlowChain) = Tangent{typeof(c)}(layers = map(elem, c.layers))
    return elem(chain)
end

function ChainRulesCore.rrule(::typeof(backward), chain::FlowChain, x::CuArray{Float32,N}, θ::CuArray{Float32,N}) where {N}
    pc = packed(chain)
    pack!(pc)                                        # Zygote differentiates the CURRENT Flux parameters
    z, ldj = _normalize(pc, x, θ, Int32(0))
    function backward_pullback(ȳ)
        z̄, j̄ = ȳ
        B = nsamples(x)
        z̄d = z̄ isa ChainRulesCore.AbstractZero ? CUDA.zeros(Float32, size(x)) : CuArray{Float32}(ChainRulesCore.unthunk(z̄))
        j̄p = j̄ isa ChainRulesCore.AbstractZero ? CU_NULL : pointer(CuArray{Float32}(ChainRulesCore.unthunk(j̄)))
        grad = CUDA.zeros(Float32, max(pc.P, 1)); x̄ = similar(x); θ̄ = similar(θ)
        ws = workspace!(pc, B)
        check(ccall((:dflow_vjp, libdflow), Cint,
                    (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Int32, CuPtr{Float32}, CuPtr{Float32},
                     CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                    pc.handle, pc.W, x, θptr(θ), B, Int32(0), z̄d, j̄p, grad, x̄, size(θ, 1) == 0 ? CU_NULL : pointer(θ̄),
                    ws, length(ws), stream()))
        return NoTangent(), tangent_of(chain, Array(grad)[1:pc.P]), x̄, θ̄
    end
    return (z, ldj), backward_pullback
end

# ---- train! on device-resident data (src/Flows.jl:380-445) -----------------------------------------------------
"η, β, ϵ of the Optimisers.Adam rule inside `Optimisers.setup(Optimisers.Adam(η), flow.model)`: first Leaf of the tree."
function adam_hyper(state)
    leaf = nothing
    walk(t) = t isa NamedTuple || t isa Tuple ? foreach(walk, t) : (hasproperty(t, :rule) && leaf === nothing && (leaf = t))
    walk(state)
    leaf === nothing && throw(ArgumentError("no Optimisers.Leaf found in the optimiser state"))
    r = leaf.rule
    return Float32(r.eta), (Float32(r.beta[1]), Float32(r.beta[2])), Float32(r.epsilon)
end

# packed Adam moments + step count, kept per optimiser-state object (the Leaf tree's own moments are left untouched)
mutable struct PackedAdam
    m::CuVector{Float32}; v::CuVector{Float32}; t::Int64
end
const _adam = IdDict{Any,PackedAdam}()
packed_adam(state, P) = get!(() -> PackedAdam(CUDA.zeros(Float32, P), CUDA.zeros(Float32, P), 0), _adam, state)

function _full_loss(pc, x, θ, idx, flags)
    acc = CUDA.zeros(Float32, 2)
    ensure_scratch!(pc, length(idx))
    check(ccall((:dflow_logpdf_sum, libdflow), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, CuPtr{Int32}, Int32, CuPtr{Float32}, Ptr{Cvoid}),
                pc.handle, pc.W, x, θptr(θ), length(idx), idx, flags, acc, stream()))
    s, bad = Array(acc)
    return bad > 0 ? NaN32 : -s / length(idx)       # loss = -mean(logpdf + ldj), src/Flows.jl:352-359
end

"Epoch order drawn on the device: out[j] = base[perm(j)] (dflow_shuffle_indices, stateless Feistel bijection)."
function device_shuffle(seed::Integer, base::CuVector{Int32})
    n = length(base); out = similar(base)
    check(ccall((:dflow_shuffle_indices, libdflow), Cint, (UInt64, Int64, Int64, Int64, CuPtr{Int32}, CuPtr{Int32}, Ptr{Cvoid}),
                UInt64(seed), n, 0, n, base, out, stream()))
    return out
end

"""
    train!(flow, data::DataArrays{Float32,2,<:CuArray,<:CuArray}, optimiser_state; epochs, batchsize, shuffle, verbose)

Same signature and epoch structure as src/Flows.jl:380-445 for data that lives on the GPU: θ is normalised in-kernel,
minibatches are gathered through the (0-based) partition indices inside the adjoint kernel, the whole epoch is enqueued by
ONE call (dflow_train_epoch: a persistent on-chip kernel for minibatches <= 512, else loss+gradient and Adam kernels per
minibatch), and the epoch-end full-set losses are dflow_logpdf_sum reductions.
"""
function train!(flow::Flow{Float32}, data::DataArrays{Float32,2,<:CuArray,<:CuArray}, optimiser_state;
                epochs::Int = 100, batchsize::Int = 64, shuffle::Bool = true, verbose::Bool = true, debug::Bool = false)
    pc = packed(flow); pack!(pc); flags = flagsof(flow)
    η, β, ϵ = adam_hyper(optimiser_state)
    st = packed_adam(optimiser_state, max(pc.P, 1))
    tr = CuArray(Int32.(collect(data.partition.training) .- 1))
    va = CuArray(Int32.(collect(data.partition.validation) .- 1))
    ws = workspace!(pc, min(batchsize, length(tr)))
    scratch = CUDA.zeros(Float32, max(pc.P, 1) + 2)
    seed0 = rand(UInt64) >> 2
    for ep in 1:epochs
        order = shuffle ? device_shuffle(seed0 + ep, tr) : tr
        t = Ref{Int64}(st.t)
        check(ccall((:dflow_train_epoch, libdflow), Cint,
                    (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Int32},
                     Int64, Int64, Float32, Float32, Float32, Float32, Ref{Int64}, Int32, CuPtr{Float32}, CuPtr{Float32},
                     CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                    pc.handle, pc.W, st.m, st.v, data.x, θptr(data.θ), order, length(order), batchsize, η, β[1], β[2], ϵ, t,
                    flags, scratch, CU_NULL, ws, length(ws), stream()))
        st.t = t[]
        push!(flow.train_loss, _full_loss(pc, data.x, data.θ, tr, flags))       # src/Flows.jl:419-421
        push!(flow.valid_loss, _full_loss(pc, data.x, data.θ, va, flags))       # src/Flows.jl:428-430
        verbose && println("epoch: $(length(flow.train_loss)) | train_loss = $(flow.train_loss[end]), valid_loss = $(flow.valid_loss[end])")
    end
    unpack!(pc)                                       # the Flux structs see the trained weights
    return debug ? (nothing, nothing) : nothing
end

# ---- one Julia process, several GPUs: dflow_dp_create_local + dflow_dp_train_step -------------------------------------
"""
    train_local!(flow, shards::Vector{<:DataArrays}, optimiser_state; devices = 0:length(shards)-1, epochs, batchsize, ...)

Data-parallel train! driven by this one process: shards[r] is resident on devices[r] (its training partition is rank r's
share of the training set).  Every minibatch is split like the shards; each device runs the adjoint on its share with the
global seed 1 / B_global, and the replicas meet in ONE fused kernel per device that sums the gradients over NVLink peer
memory and applies Adam (bit-identical replicas, no NCCL call, no host round trip per step).
"""
function train_local!(flow::Flow{Float32}, shards::Vector, optimiser_state; devices = collect(0:length(shards)-1),
                      epochs::Int = 100, batchsize::Int = 64, shuffle::Bool = true, verbose::Bool = true)
    R = length(shards); flags = flagsof(flow)
    η, β, ϵ = adam_hyper(optimiser_state)
    reps = PackedChain[]; ms = CuVector{Float32}[]; vs = CuVector{Float32}[]; l2 = CuVector{Float32}[]; trs = CuVector{Int32}[]
    for r in 1:R
        CUDA.device!(devices[r])
        pc = PackedChain(flow.model; θ_min = flow.metadata.θ_min, θ_max = flow.metadata.θ_max)
        push!(reps, pc); push!(ms, CUDA.zeros(Float32, max(pc.P, 1))); push!(vs, CUDA.zeros(Float32, max(pc.P, 1)))
        push!(l2, CUDA.zeros(Float32, 2)); push!(trs, CuArray(Int32.(collect(shards[r].partition.training) .- 1)))
    end
    P = max(reps[1].P, 1)
    dps = fill(C_NULL, R)
    check(ccall((:dflow_dp_create_local, libdflow), Cint, (Int32, Ptr{Int32}, Int64, Ptr{Ptr{Cvoid}}), R, Int32.(devices), P, dps))
    ntr = sum(length, trs); t = 0; seed0 = rand(UInt64) >> 2
    share(nb, r) = (per = cld(nb, R); lo = min((r - 1) * per, nb); min(lo + per, nb) - lo)   # shard_range of flows.py
    for ep in 1:epochs
        orders = [(CUDA.device!(devices[r]); shuffle ? device_shuffle(seed0 + ep + 1000003r, trs[r]) : trs[r]) for r in 1:R]
        cur = zeros(Int, R)
        for b0 in 0:batchsize:ntr-1
            nb = min(batchsize, ntr - b0)
            ks = [min(share(nb, r), length(orders[r]) - cur[r]) for r in 1:R]
            t += 1
            sh = [begin
                      CUDA.device!(devices[r]); ws = workspace!(reps[r], max(ks[r], 1))
                      DpShard(reps[r].handle, pointer(reps[r].W), pointer(ms[r]), pointer(vs[r]), pointer(shards[r].x),
                              θptr(shards[r].θ), ks[r], pointer(orders[r]) + 4 * cur[r], pointer(ws), length(ws), pointer(l2[r]), C_NULL)
                  end for r in 1:R]
            check(ccall((:dflow_dp_train_step, libdflow), Cint,
                        (Ptr{Ptr{Cvoid}}, Int32, Ptr{DpShard}, Float32, Int32, Float32, Float32, Float32, Float32, Int64),
                        dps, R, sh, 1f0 / sum(ks), flags, η, β[1], β[2], ϵ, t))
            cur .+= ks
        end
        rc = ccall((:dflow_dp_sync, libdflow), Cint, (Ptr{Ptr{Cvoid}}, Int32, Ptr{DpShard}), dps, R, C_NULL)
        rc > 0 && error("data-parallel step: a device missed the peer barrier; its update was skipped")
        check(min(rc, 0))
        verbose && println("epoch: $ep done ($(t) steps)")
    end
    foreach(dp -> ccall((:dflow_dp_destroy, libdflow), Cint, (Ptr{Cvoid},), dp), dps)
    CUDA.device!(devices[1]); unpack!(reps[1])
    return nothing
end

# tuning knobs (dflow_set_tuning): "tc_mode" 1 / 0 / -1 forces an eligible hidden <= 64 chain onto / off the tensor cores,
# "tc_ws_budget_mb" caps the adjoint workspace, "epoch_kernel" -1 keeps train! on per-minibatch launches.
tune!(pc, key::AbstractString, value::Integer) =
    check(ccall((:dflow_set_tuning, libdflow), Cint, (Ptr{Cvoid}, Cstring, Int32), pc.handle, key, value))

end # module
