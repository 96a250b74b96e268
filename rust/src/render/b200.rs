//! UNVERIFIED (never compiled here: no cargo/rustc in the image).
//! `B200Renderer`: the reference's `Renderer` + `GraphWatcher` traits (src/render/renderer.rs:6-17,
//! src/routing/graphwatcher.rs:4-9) over the C ABI of include/friendship_b200.h.
//! It mirrors what `RefRenderer` does with its arguments (src/render/reference.rs:98-137): nested effects are
//! walked with `iter_nodes()/iter_edges()` and handed over as definitions, children first.

use std::collections::HashMap;
use std::os::raw::{c_char, c_int};
use std::rc::Rc;

use jagged_array::Jagged2;
use ndarray::Array2;
use streaming_iterator::StreamingIterator;

use render::Renderer;
use routing::{Edge, Effect, GraphWatcher, NodeData, NodeHandle};
use routing::effect::{EffectData, PrimitiveEffect};

#[repr(C)] #[derive(Copy, Clone)]
struct FrbEdge { from: u32, to: u32, from_slot: u32, to_slot: u32 }
#[repr(C)] #[derive(Copy, Clone)]
struct FrbNode { handle: u32, kind: u32, key: u64 }
#[repr(C)]
struct FrbConfig { device: i32, flags: u32, osc_anchor: u32, n_devices: u32 }
enum FrbRenderer {}

extern "C" {
    fn frb_create(cfg: *const FrbConfig) -> *mut FrbRenderer;
    fn frb_destroy(r: *mut FrbRenderer);
    fn frb_last_error(r: *const FrbRenderer) -> *const c_char;
    fn frb_define_effect(r: *mut FrbRenderer, key: u64, nodes: *const FrbNode, n_nodes: u32,
                         edges: *const FrbEdge, n_edges: u32) -> c_int;
    fn frb_add_node(r: *mut FrbRenderer, handle: u32, kind: u32, key: u64) -> c_int;
    fn frb_del_node(r: *mut FrbRenderer, handle: u32) -> c_int;
    fn frb_add_edge(r: *mut FrbRenderer, e: FrbEdge) -> c_int;
    fn frb_del_edge(r: *mut FrbRenderer, e: FrbEdge) -> c_int;
    fn frb_fill_buffer(r: *mut FrbRenderer, out: *mut f32, n_slots: u32, n_times: u64, idx: u64,
                       in_data: *const f32, in_row_offsets: *const u64, n_in_rows: u32) -> c_int;
}

const KIND_EFFECT: u32 = 16;

fn prim_kind(p: PrimitiveEffect) -> u32 {
    match p {
        PrimitiveEffect::Delay => 0, PrimitiveEffect::F32Constant => 1, PrimitiveEffect::Sum2 => 2,
        PrimitiveEffect::Multiply => 3, PrimitiveEffect::Divide => 4, PrimitiveEffect::Modulo => 5,
        PrimitiveEffect::Minimum => 6,
    }
}
fn handle_u32(h: &NodeHandle) -> u32 { h.node_handle().get().unwrap_or(0) }
fn edge_c(e: &Edge) -> FrbEdge {
    FrbEdge { from: handle_u32(&e.from_full()), to: handle_u32(&e.to_full()), from_slot: e.from_slot(), to_slot: e.to_slot() }
}

pub struct B200Renderer {
    raw: *mut FrbRenderer,
    /// definition key per effect already handed to the library (keyed by the Rc's address)
    keys: HashMap<*const Effect, u64>,
    next_key: u64,
}

impl Default for B200Renderer {
    fn default() -> Self {
        // FRB_N_DEVICES=8: the eight B200s of a box behind this one renderer (voices sharded v mod N inside the library);
        // `Dispatch<B200Renderer, C>` and its callers do not change
        let n_devices = std::env::var("FRB_N_DEVICES").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let cfg = FrbConfig { device: 0, flags: 0, osc_anchor: 0, n_devices };
        let raw = unsafe { frb_create(&cfg) };
        assert!(!raw.is_null(), "frb_create failed: no CUDA device (there is no CPU fallback)");
        B200Renderer { raw, keys: HashMap::new(), next_key: 1 }
    }
}
impl Drop for B200Renderer { fn drop(&mut self) { unsafe { frb_destroy(self.raw) } } }

impl B200Renderer {
    fn check(&self, rc: c_int) {
        // the trait methods return (); the reference panics on broken invariants (reference.rs:69,71,131,145,199)
        if rc != 0 {
            let msg = unsafe { std::ffi::CStr::from_ptr(frb_last_error(self.raw)) };
            panic!("friendship_b200: [{}] {}", rc, msg.to_string_lossy());
        }
    }
    /// Returns (kind, key) for a node's data, defining nested effects (children first) on the way.
    fn kind_of(&mut self, data: &NodeData) -> (u32, u64) {
        match *data.data() {
            EffectData::Primitive(p) => (prim_kind(p), 0),
            EffectData::RouteGraph(ref graph) => {
                let ptr = Rc::as_ptr(data);
                if let Some(k) = self.keys.get(&ptr) { return (KIND_EFFECT, *k); }
                let mut nodes = vec![];
                for (hnd, child) in graph.iter_nodes() {
                    let (kind, key) = self.kind_of(child);
                    nodes.push(FrbNode { handle: handle_u32(hnd), kind, key });
                }
                let edges: Vec<FrbEdge> = graph.iter_edges().map(edge_c).collect();
                let key = self.next_key; self.next_key += 1;
                let rc = unsafe { frb_define_effect(self.raw, key, nodes.as_ptr(), nodes.len() as u32,
                                                    edges.as_ptr(), edges.len() as u32) };
                self.check(rc);
                self.keys.insert(ptr, key);
                (KIND_EFFECT, key)
            }
        }
    }
}

impl GraphWatcher for B200Renderer {
    fn on_add_node(&mut self, handle: &NodeHandle, data: &NodeData) {
        let (kind, key) = self.kind_of(data);
        let rc = unsafe { frb_add_node(self.raw, handle_u32(handle), kind, key) }; self.check(rc);
    }
    fn on_del_node(&mut self, handle: &NodeHandle) {
        let rc = unsafe { frb_del_node(self.raw, handle_u32(handle)) }; self.check(rc);
    }
    fn on_add_edge(&mut self, edge: &Edge) {
        let rc = unsafe { frb_add_edge(self.raw, edge_c(edge)) }; self.check(rc);
    }
    fn on_del_edge(&mut self, edge: &Edge) {
        let rc = unsafe { frb_del_edge(self.raw, edge_c(edge)) }; self.check(rc);
    }
}

impl Renderer for B200Renderer {
    fn fill_buffer(&mut self, buff: &mut Array2<f32>, idx: u64, inputs: Jagged2<f32>) {
        let (n_slots, n_times) = buff.dim();
        // Jagged2 -> (flat data, row offsets)
        let mut data: Vec<f32> = vec![];
        let mut offs: Vec<u64> = vec![0];
        let mut stream = inputs.stream();
        while let Some(row) = stream.next() {
            data.extend_from_slice(row);
            offs.push(data.len() as u64);
        }
        let out = buff.as_slice_mut().expect("Dispatch allocates a C-contiguous buffer (dispatch.rs:149)");
        let rc = unsafe { frb_fill_buffer(self.raw, out.as_mut_ptr(), n_slots as u32, n_times as u64, idx,
                                          data.as_ptr(), offs.as_ptr(), (offs.len() - 1) as u32) };
        self.check(rc);
    }
}
