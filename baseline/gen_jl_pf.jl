# baseline/gen_jl_pf.jl -- the BASELINE.json configs as plain Gen.jl programs (reference API only).
# Not runnable in this repository's environment (no Julia); kept for machines that have Julia + Gen:
#     julia --threads 1 baseline/gen_jl_pf.jl bearings 100000 50
using Gen, Random

struct State; x::Float64; y::Float64; vx::Float64; vy::Float64; end

# cfg 3: the tutorial's bearings-only tracker (docs/src/tutorials/smc.md:607-661)
@gen (static) function bearings_kernel(t::Int, prev::State, velocity_var::Float64, measurement_noise::Float64)
    vx ~ normal(prev.vx, sqrt(velocity_var))
    vy ~ normal(prev.vy, sqrt(velocity_var))
    x = prev.x + vx
    y = prev.y + vy
    z ~ normal(atan(y, x), measurement_noise)
    return State(x, y, vx, vy)
end
bearings_chain = Unfold(bearings_kernel)
@gen (static) function bearings_model(T::Int)
    x0 ~ normal(0.01, 0.01); y0 ~ normal(0.95, 0.01); vx0 ~ normal(0.002, 0.01); vy0 ~ normal(-0.013, 0.01)
    z0 ~ normal(atan(y0, x0), 0.005)
    chain ~ bearings_chain(T, State(x0, y0, vx0, vy0), 1e-6, 0.005)
end

# cfg 1: 1-D linear-Gaussian SSM
@gen (static) function lgssm_kernel(t::Int, x_prev::Float64, a::Float64, q::Float64, r::Float64)
    x ~ normal(a * x_prev, q)
    y ~ normal(x, r)
    return x
end
lgssm_chain = Unfold(lgssm_kernel)
@gen (static) function lgssm_model(T::Int)
    x0 ~ normal(0.0, 1.0)
    y0 ~ normal(x0, 1.0)
    chain ~ lgssm_chain(T, x0, 0.9, 0.5, 1.0)
end

# cfg 2: stochastic volatility
@gen (static) function sv_kernel(t::Int, h_prev::Float64, mu::Float64, phi::Float64, sigma::Float64)
    h ~ normal(mu + phi * (h_prev - mu), sigma)
    y ~ normal(0.0, exp(h / 2))
    return h
end
sv_chain = Unfold(sv_kernel)
@gen (static) function sv_model(T::Int)
    h0 ~ normal(-1.0, 0.25 / sqrt(1 - 0.95^2))
    y0 ~ normal(0.0, exp(h0 / 2))
    chain ~ sv_chain(T, h0, -1.0, 0.95, 0.25)
end

Gen.@load_generated_functions

function run_pf(model, obs0_addr, obs_addr, ys, N)
    state = initialize_particle_filter(model, (0,), choicemap((obs0_addr, ys[1])), N)
    for t in 1:length(ys)-1
        maybe_resample!(state, ess_threshold=N / 2)
        particle_filter_step!(state, (t,), (UnknownChange(),), choicemap((:chain => t => obs_addr, ys[t+1])))
    end
    log_ml_estimate(state)
end

function main()
    which, N, T = ARGS[1], parse(Int, ARGS[2]), parse(Int, ARGS[3])
    Random.seed!(1)
    ys = randn(T + 1)   # replace by the series of tests/fixtures.py for a like-for-like comparison
    model, a0, a = which == "bearings" ? (bearings_model, :z0, :z) : which == "sv" ? (sv_model, :y0, :y) : (lgssm_model, :y0, :y)
    run_pf(model, a0, a, ys[1:3], 100)  # compile
    t = @elapsed lml = run_pf(model, a0, a, ys, N)
    println("$(which): N=$N T=$T threads=$(Threads.nthreads()) $(N * T / t) particle-steps/s, log ML estimate $lml")
end
main()
