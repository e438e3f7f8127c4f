//! ndarray 0.17.2 on the host cores, for the five BASELINE configs.  Prints one JSON object per case with
//! the same algorithmic byte / flop counts bench.py uses, so the numbers sit beside `profiles/r0N_bench_*.json`.
//!
//!   cargo run --release -- c1 | c2 | c3 | c4 | c5 | all   (RAYON_NUM_THREADS selects the thread count)
//!
//! ndarray has no parallel `sum_axis` / `var_axis` / `fold_axis`: those are timed serial, as a user of the crate gets them;
//! the global sum uses the documented parallel idiom (`axis_iter().into_par_iter().map(sum).sum()`, tests/par_rayon.rs:19).
use ndarray::parallel::prelude::*;
use ndarray::prelude::*;
use ndarray::{linalg::general_mat_mul, s, Zip};
use std::time::Instant;

fn lcg(seed: u64, n: usize) -> Vec<f32> {
    let mut x = seed;
    (0..n)
        .map(|_| {
            x = x.wrapping_mul(6364136223846793005).wrapping_add(1442695040888963407);
            ((x >> 40) as f32 / (1u64 << 24) as f32) * 2.0 - 1.0
        })
        .collect()
}

fn time<F: FnMut()>(mut f: F, reps: usize) -> f64 {
    f();
    let mut best = f64::INFINITY;
    for _ in 0..reps {
        let t = Instant::now();
        f();
        best = best.min(t.elapsed().as_secs_f64());
    }
    best
}

fn report(case: &str, bytes: f64, secs: f64) {
    println!(
        "{{\"impl\": \"ndarray-0.17.2\", \"case\": \"{case}\", \"GB/s\": {:.2}, \"ms\": {:.3}, \"threads\": {}}}",
        bytes / secs / 1e9,
        secs * 1e3,
        rayon::current_num_threads()
    );
}

fn c1() {
    let n = 4096;
    let mut a = Array2::from_shape_vec((n, n), lcg(1, n * n)).unwrap();
    let b = Array2::from_shape_vec((n, n), lcg(2, n * n)).unwrap();
    let c = Array2::from_shape_vec((n, n), lcg(3, n * n)).unwrap();
    let t = time(|| Zip::from(&mut a).and(&b).and(&c).par_for_each(|a, &b, &c| *a = b * c + *a), 10);
    report("C1_zip_a=b*c+a_f32_4096^2", 4.0 * 4.0 * (n * n) as f64, t);
    for ax in 0..2 {
        let t = time(|| { std::hint::black_box(a.sum_axis(Axis(ax))); }, 10);
        report(&format!("C1_sum_axis{ax}_f32_4096^2"), 4.0 * (n * n + n) as f64, t);
    }
}

fn c2() {
    let n = 512;
    let x = Array3::from_shape_vec((n, n, n), lcg(1, n * n * n)).unwrap();
    let y = Array3::from_shape_vec((n, n, n), lcg(2, n * n * n)).unwrap();
    let row = Array1::from_vec(lcg(3, n));
    let mut out = Array3::<f32>::zeros((n, n, n));
    let el = (n * n * n) as f64;
    let t1 = time(|| Zip::from(&mut out).and(x.t()).and(&y).par_for_each(|o, &a, &b| *o = a + b), 5);
    report("C2_transposed_x.t()+y", 3.0 * 4.0 * el, t1);
    let t2 = time(
        || {
            Zip::from(out.slice_mut(s![..n / 2, .., ..]))
                .and(x.slice(s![..;-2, .., ..]))
                .and(y.slice(s![..;2, .., ..]))
                .par_for_each(|o, &a, &b| *o = a + b)
        },
        5,
    );
    report("C2_neg_step_s![..;-2]+s![..;2]", 3.0 * 4.0 * el / 2.0, t2);
    let t3 = time(|| Zip::from(&mut out).and(&x).and_broadcast(&row).par_for_each(|o, &a, &r| *o = a * r), 5);
    report("C2_broadcast_row_x*row", 2.0 * 4.0 * el + 4.0 * n as f64, t3);
    // the bench.py metric: all three Zips as one step
    let bytes = 3.0 * 4.0 * el + 3.0 * 4.0 * el / 2.0 + 2.0 * 4.0 * el + 4.0 * n as f64;
    report("C2_step_zip_views_hbm_gbps", bytes, t1 + t2 + t3);
}

fn c3() {
    let n = 16384;
    let mut x = 88172645463325252u64;
    let f = Array2::from_shape_fn((n, n), |_| {
        x ^= x << 13;
        x ^= x >> 7;
        x ^= x << 17;
        (x >> 11) as f64 / (1u64 << 53) as f64 - 0.5
    });
    let bytes = 8.0 * (n * n + n) as f64;
    for ax in 0..2 {
        let t = time(|| { std::hint::black_box(f.sum_axis(Axis(ax))); }, 3);
        report(&format!("C3_sum_axis{ax}_f64_16384^2"), bytes, t);
        let t = time(|| { std::hint::black_box(f.mean_axis(Axis(ax))); }, 3);
        report(&format!("C3_mean_axis{ax}_f64_16384^2"), bytes, t);
        let t = time(|| { std::hint::black_box(f.var_axis(Axis(ax), 1.0)); }, 1);
        report(&format!("C3_var_axis{ax}_f64_16384^2"), bytes, t);
    }
    let t = time(|| { std::hint::black_box(f.axis_iter(Axis(0)).into_par_iter().map(|r| r.sum()).sum::<f64>()); }, 3);
    report("C3_sum_f64_16384^2_par_rows", 8.0 * (n * n) as f64, t);
    let i = f.mapv(|v| (v * 1e18) as i64);
    for ax in 0..2 {
        let t = time(|| { std::hint::black_box(i.fold_axis(Axis(ax), i64::MAX, |&m, &v| m.min(v))); }, 3);
        report(&format!("C3_fold_axis{ax}_min_i64_16384^2"), bytes, t);
    }
}

fn c4() {
    let n = 8192;
    let a64 = Array2::from_shape_fn((n, n), |(i, j)| ((i * 31 + j * 17) % 101) as f64 / 101.0 - 0.5);
    let b64 = a64.t().to_owned();
    let mut c64 = Array2::<f64>::zeros((n, n));
    let t = time(|| general_mat_mul(1.0, &a64, &b64, 0.0, &mut c64), 1);
    println!("{{\"impl\": \"ndarray-0.17.2\", \"case\": \"C4_gemm_f64_8192^3\", \"TFLOP/s\": {:.3}, \"ms\": {:.1}}}", 2.0 * (n as f64).powi(3) / t / 1e12, t * 1e3);
    let a32 = a64.mapv(|v| v as f32);
    let b32 = b64.mapv(|v| v as f32);
    let mut c32 = Array2::<f32>::zeros((n, n));
    let t = time(|| general_mat_mul(1.0, &a32, &b32, 0.0, &mut c32), 1);
    println!("{{\"impl\": \"ndarray-0.17.2\", \"case\": \"C4_gemm_f32_8192^3\", \"TFLOP/s\": {:.3}, \"ms\": {:.1}}}", 2.0 * (n as f64).powi(3) / t / 1e12, t * 1e3);
}

/// splitmix64 output `i` of a generator seeded with `seed`, top 24 bits -> [-1, 1): the harness's index-addressable generator
/// (ndarray_b200/csrc/rng.cu, oracle orc_counter_fill), so this program sees the same data as bench.py.
fn counter_f32(seed: u64, i: u64) -> f32 {
    let mut z = seed.wrapping_add((i + 1).wrapping_mul(0x9E3779B97F4A7C15));
    z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
    z ^= z >> 31;
    ((z >> 40) as u32 as f32) * (1.0 / 8388608.0) - 1.0
}

/// Config 5, the round-2 bench headline: f32 [rows, 65536] `a = b*c + a` (par_for_each), then the parallel global sum idiom of
/// tests/par_rayon.rs:19.  `rows` = 8192 is the bounded sample bench.py's CPU arm uses (1/8 of the 2^32-element problem);
/// pass 65536 as the second argument for the whole thing (48 GiB of host memory).
fn c5(rows: usize) {
    let cols = 65536usize;
    let fill = |seed: u64| Array2::from_shape_fn((rows, cols), |(i, j)| counter_f32(seed, (i * cols + j) as u64));
    let (mut a, b, c) = (fill(0x5EED0001), fill(0x5EED0002), fill(0x5EED0003));
    let mut total = 0f32;
    let t = time(
        || {
            Zip::from(&mut a).and(&b).and(&c).par_for_each(|a, &b, &c| *a = b * c + *a);
            total = a.axis_iter(Axis(0)).into_par_iter().map(|r| r.sum()).sum::<f32>();
        },
        10,
    );
    std::hint::black_box(total);
    report(&format!("C5_zip_global_sum_hbm_gbps_f32_{rows}x{cols}"), 20.0 * (rows * cols) as f64, t);
}

fn main() {
    let which = std::env::args().nth(1).unwrap_or_else(|| "all".into());
    if which == "c1" || which == "all" { c1(); }
    if which == "c2" || which == "all" { c2(); }
    if which == "c3" || which == "all" { c3(); }
    if which == "c4" || which == "all" { c4(); }
    if which == "c5" || which == "all" {
        c5(std::env::args().nth(2).and_then(|s| s.parse().ok()).unwrap_or(8192));
    }
}
