//! A recording `EvalAtRow`: running a `FrameworkEval::evaluate` over it (the reference's `MachineEval<C>`,
//! prover/src/components/mod.rs:48-57, and every extension's eval, prover/src/extensions/*) produces the AIR bytecode that
//! `nb200_air_load` executes on the GPU — constraint program, mask table, logup fractions and batching.
//!
//! `nexus_zkvm_b200/air.py` is the executable specification: same node kinds, same common-subexpression table, same
//! column allocation order as `TraceLocationAllocator::default()` (machine.rs:265), same logup constraint shapes as
//! stwo-constraint-framework's `logup_proxy!` (cumulative-sum columns, [-1, 0] mask on the last one, `cumsum_shift`),
//! same emission (post-order per sink, leaves rematerialised, registers reused after the last consumer).
//! The format ('NBAR' v1) is documented at the top of air.py.
use std::cell::RefCell;
use std::collections::HashMap;
use std::ops::{Add, AddAssign, Mul, Neg, Sub};
use std::rc::Rc;

use num_traits::{One, Zero};
use stwo::core::fields::m31::BaseField;
use stwo::core::fields::qm31::SecureField;
use stwo::core::fields::FieldExpOps;
use stwo::core::lookups::utils::Fraction;
use stwo_constraint_framework::preprocessed_columns::PreProcessedColumnId;
use stwo_constraint_framework::{Batching, EvalAtRow, Relation, RelationEntry, INTERACTION_TRACE_IDX, ORIGINAL_TRACE_IDX, PREPROCESSED_TRACE_IDX};

const P: u32 = (1 << 31) - 1;
pub const NO_PARAM: u32 = 0xFFFF_FFFF;

// opcodes (air.py)
const OP_LOADM: u32 = 0; const OP_CONSTB: u32 = 1; const OP_ADDB: u32 = 2; const OP_SUBB: u32 = 3; const OP_MULB: u32 = 4; const OP_NEGB: u32 = 5;
const OP_PARAME: u32 = 6; const OP_ADDE: u32 = 8; const OP_SUBE: u32 = 9; const OP_MULE: u32 = 10; const OP_NEGE: u32 = 11;
const OP_ADDEB: u32 = 12; const OP_SUBEB: u32 = 13; const OP_MULEB: u32 = 14; const OP_BTOE: u32 = 15; const OP_LOADME: u32 = 16;
const OP_CONSTRB: u32 = 17; const OP_CONSTRE: u32 = 18; const OP_FRAC: u32 = 19;

#[derive(Clone, Copy, PartialEq, Eq, Hash, Debug)]
enum Kind { B, E }
#[derive(Clone, Copy, PartialEq, Eq, Hash, Debug)]
enum Op { Mask, MaskE, Const, Param, Add, Sub, Mul, Neg, AddEb, SubEb, MulEb, BtoE }
#[derive(Clone, Copy, PartialEq, Eq, Hash, Debug)]
struct Node { op: Op, a: u32, b: u32, kind: Kind }

/// Proof-wide parameter table shared by all components of a machine: lookup elements (z, alpha^i) of every relation, one
/// `cumsum_shift` per component (= claimed_sum / 2^log_size, filled by the host after interaction-trace generation) and the
/// rare secure-field literals of an AIR.  Mirrors `air.py:Air.alloc_param`.
#[derive(Default)]
pub struct ParamTable {
    pub names: Vec<String>,
    /// parameters whose value is a compile-time secure-field literal (index, value)
    pub literals: Vec<(u32, SecureField)>,
    relations: HashMap<String, RelationParams>,
    literal_index: HashMap<[u32; 4], u32>,
}
#[derive(Clone, Debug)]
pub struct RelationParams { pub z: u32, pub alpha_powers: Vec<u32> }
impl ParamTable {
    pub fn alloc(&mut self, name: &str) -> u32 { self.names.push(name.to_string()); (self.names.len() - 1) as u32 }
    /// `relation!(Name, N)`: z, alpha^0 .. alpha^(N-1); idempotent per name
    pub fn relation(&mut self, name: &str, size: usize) -> RelationParams {
        if let Some(r) = self.relations.get(name) { return r.clone(); }
        let z = self.alloc(&format!("{name}.z"));
        let alpha_powers = (0..size).map(|i| self.alloc(&format!("{name}.alpha^{i}"))).collect();
        let r = RelationParams { z, alpha_powers };
        self.relations.insert(name.to_string(), r.clone());
        r
    }
}

/// `TraceLocationAllocator`: next free column per tree + the preprocessed column ids seen so far
#[derive(Default)]
pub struct Allocator { pub next_col: [u32; 3], pub preprocessed: HashMap<String, u32> }

struct Graph {
    nodes: Vec<Node>,
    cse: HashMap<Node, u32>,
    masks: Vec<(u32, u32, i32)>,
    mask_slot: HashMap<(u32, u32, i32), u32>,
    constraints: Vec<u32>,
    fracs: Vec<(u32, u32)>,
    batching: Option<Vec<usize>>,
    cumsum_shift_param: u32,
    interaction_col0: Option<u32>,
    alloc: Rc<RefCell<Allocator>>,
    params: Rc<RefCell<ParamTable>>,
}
impl Graph {
    fn node(&mut self, op: Op, a: u32, b: u32, kind: Kind) -> u32 {
        let n = Node { op, a, b, kind };
        if let Some(&id) = self.cse.get(&n) { return id; }
        self.nodes.push(n);
        let id = (self.nodes.len() - 1) as u32;
        self.cse.insert(n, id);
        id
    }
    fn slot(&mut self, tree: u32, col: u32, off: i32) -> u32 {
        if let Some(&s) = self.mask_slot.get(&(tree, col, off)) { return s; }
        self.masks.push((tree, col, off));
        let s = (self.masks.len() - 1) as u32;
        self.mask_slot.insert((tree, col, off), s);
        s
    }
    fn konst(&mut self, v: BaseField) -> u32 { self.node(Op::Const, v.0 % P, 0, Kind::B) }
    fn secure_literal(&mut self, v: SecureField) -> u32 {
        let [c0, c1, c2, c3] = v.to_m31_array();
        if c1.0 == 0 && c2.0 == 0 && c3.0 == 0 { let b = self.konst(c0); return self.node(Op::BtoE, b, 0, Kind::E); }
        // lookup elements reach the evaluator as SecureField literals (they are drawn before the component is built, machine.rs:239-240 and
        // MachineEval::new): each distinct literal becomes one slot of the parameter table, so the PROGRAM does not depend on the drawn values
        let key = [c0.0, c1.0, c2.0, c3.0];
        let idx = { let mut p = self.params.borrow_mut();
                    if let Some(&i) = p.literal_index.get(&key) { i } else { let i = p.alloc("literal"); p.literals.push((i, v)); p.literal_index.insert(key, i); i } };
        self.node(Op::Param, idx, 0, Kind::E)
    }
}
type G = Rc<RefCell<Graph>>;

/// `EvalAtRow::F`: a base-field value of the row evaluation — a literal or a node of the recording graph
#[derive(Clone, Debug)]
pub enum BaseExpr { Lit(BaseField), Node(GraphRef, u32) }
/// `EvalAtRow::EF`
#[derive(Clone, Debug)]
pub enum ExtExpr { Lit(SecureField), Node(GraphRef, u32) }
#[derive(Clone)]
pub struct GraphRef(G);
impl std::fmt::Debug for GraphRef { fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result { write!(f, "<graph>") } }

impl BaseExpr {
    fn id_in(&self, g: &G) -> u32 { match self { BaseExpr::Lit(v) => g.borrow_mut().konst(*v), BaseExpr::Node(_, id) => *id } }
    fn graph(&self) -> Option<G> { if let BaseExpr::Node(g, _) = self { Some(g.0.clone()) } else { None } }
}
impl ExtExpr {
    fn id_in(&self, g: &G) -> u32 { match self { ExtExpr::Lit(v) => g.borrow_mut().secure_literal(*v), ExtExpr::Node(_, id) => *id } }
    fn graph(&self) -> Option<G> { if let ExtExpr::Node(g, _) = self { Some(g.0.clone()) } else { None } }
}
fn bin_b(op: Op, x: &BaseExpr, y: &BaseExpr, lit: fn(BaseField, BaseField) -> BaseField) -> BaseExpr {
    match x.graph().or_else(|| y.graph()) {
        None => match (x, y) { (BaseExpr::Lit(a), BaseExpr::Lit(b)) => BaseExpr::Lit(lit(*a, *b)), _ => unreachable!() },
        Some(g) => { let (a, b) = (x.id_in(&g), y.id_in(&g)); let id = g.borrow_mut().node(op, a, b, Kind::B); BaseExpr::Node(GraphRef(g), id) }
    }
}
fn bin_e(op: Op, x: &ExtExpr, y: &ExtExpr, lit: fn(SecureField, SecureField) -> SecureField) -> ExtExpr {
    match x.graph().or_else(|| y.graph()) {
        None => match (x, y) { (ExtExpr::Lit(a), ExtExpr::Lit(b)) => ExtExpr::Lit(lit(*a, *b)), _ => unreachable!() },
        Some(g) => { let (a, b) = (x.id_in(&g), y.id_in(&g)); let id = g.borrow_mut().node(op, a, b, Kind::E); ExtExpr::Node(GraphRef(g), id) }
    }
}
/// E (op) B with the `_eb` node kinds of air.py (`B - E` is lifted: btoe(B) - E)
fn bin_eb(op: Op, x: &ExtExpr, y: &BaseExpr, lit: fn(SecureField, BaseField) -> SecureField) -> ExtExpr {
    match x.graph().or_else(|| y.graph()) {
        None => match (x, y) { (ExtExpr::Lit(a), BaseExpr::Lit(b)) => ExtExpr::Lit(lit(*a, *b)), _ => unreachable!() },
        Some(g) => { let (a, b) = (x.id_in(&g), y.id_in(&g)); let id = g.borrow_mut().node(op, a, b, Kind::E); ExtExpr::Node(GraphRef(g), id) }
    }
}
fn to_e(x: &BaseExpr) -> ExtExpr {
    match x { BaseExpr::Lit(v) => ExtExpr::Lit(SecureField::from(*v)),
              BaseExpr::Node(g, id) => { let n = g.0.borrow_mut().node(Op::BtoE, *id, 0, Kind::E); ExtExpr::Node(g.clone(), n) } }
}

// ---- operator surface required by `EvalAtRow::{F, EF}` (stwo constraint-framework/src/lib.rs) ----
macro_rules! impl_bb { ($tr:ident, $f:ident, $op:expr, $lit:expr) => {
    impl $tr<BaseExpr> for BaseExpr { type Output = BaseExpr; fn $f(self, o: BaseExpr) -> BaseExpr { bin_b($op, &self, &o, $lit) } }
    impl $tr<BaseField> for BaseExpr { type Output = BaseExpr; fn $f(self, o: BaseField) -> BaseExpr { bin_b($op, &self, &BaseExpr::Lit(o), $lit) } }
} }
impl_bb!(Add, add, Op::Add, |a, b| a + b);
impl_bb!(Sub, sub, Op::Sub, |a, b| a - b);
impl_bb!(Mul, mul, Op::Mul, |a, b| a * b);
macro_rules! impl_ee { ($tr:ident, $f:ident, $op:expr, $opeb:expr, $lit:expr, $liteb:expr) => {
    impl $tr<ExtExpr> for ExtExpr { type Output = ExtExpr; fn $f(self, o: ExtExpr) -> ExtExpr { bin_e($op, &self, &o, $lit) } }
    impl $tr<SecureField> for ExtExpr { type Output = ExtExpr; fn $f(self, o: SecureField) -> ExtExpr { bin_e($op, &self, &ExtExpr::Lit(o), $lit) } }
    impl $tr<BaseExpr> for ExtExpr { type Output = ExtExpr; fn $f(self, o: BaseExpr) -> ExtExpr { bin_eb($opeb, &self, &o, $liteb) } }
    impl $tr<BaseField> for ExtExpr { type Output = ExtExpr; fn $f(self, o: BaseField) -> ExtExpr { bin_eb($opeb, &self, &BaseExpr::Lit(o), $liteb) } }
} }
impl_ee!(Add, add, Op::Add, Op::AddEb, |a, b| a + b, |a, b| a + b);
impl_ee!(Sub, sub, Op::Sub, Op::SubEb, |a, b| a - b, |a, b| a - b);
impl_ee!(Mul, mul, Op::Mul, Op::MulEb, |a, b| a * b, |a, b| a * b);
// F (op) SecureField -> EF
impl Add<SecureField> for BaseExpr { type Output = ExtExpr; fn add(self, o: SecureField) -> ExtExpr { bin_eb(Op::AddEb, &ExtExpr::Lit(o), &self, |a, b| a + b) } }
impl Mul<SecureField> for BaseExpr { type Output = ExtExpr; fn mul(self, o: SecureField) -> ExtExpr { bin_eb(Op::MulEb, &ExtExpr::Lit(o), &self, |a, b| a * b) } }
impl Neg for BaseExpr { type Output = BaseExpr;
    fn neg(self) -> BaseExpr { match self { BaseExpr::Lit(v) => BaseExpr::Lit(-v), BaseExpr::Node(g, id) => { let n = g.0.borrow_mut().node(Op::Neg, id, 0, Kind::B); BaseExpr::Node(g, n) } } } }
impl Neg for ExtExpr { type Output = ExtExpr;
    fn neg(self) -> ExtExpr { match self { ExtExpr::Lit(v) => ExtExpr::Lit(-v), ExtExpr::Node(g, id) => { let n = g.0.borrow_mut().node(Op::Neg, id, 0, Kind::E); ExtExpr::Node(g, n) } } } }
impl AddAssign<BaseExpr> for BaseExpr { fn add_assign(&mut self, o: BaseExpr) { *self = self.clone() + o } }
impl AddAssign<BaseField> for BaseExpr { fn add_assign(&mut self, o: BaseField) { *self = self.clone() + o } }
impl AddAssign<ExtExpr> for ExtExpr { fn add_assign(&mut self, o: ExtExpr) { *self = self.clone() + o } }
impl From<BaseField> for BaseExpr { fn from(v: BaseField) -> Self { BaseExpr::Lit(v) } }
impl From<SecureField> for ExtExpr { fn from(v: SecureField) -> Self { ExtExpr::Lit(v) } }
impl From<BaseExpr> for ExtExpr { fn from(v: BaseExpr) -> Self { to_e(&v) } }
impl Zero for BaseExpr { fn zero() -> Self { BaseExpr::Lit(BaseField::zero()) } fn is_zero(&self) -> bool { matches!(self, BaseExpr::Lit(v) if v.is_zero()) } }
impl One for BaseExpr { fn one() -> Self { BaseExpr::Lit(BaseField::one()) } }
impl Zero for ExtExpr { fn zero() -> Self { ExtExpr::Lit(SecureField::zero()) } fn is_zero(&self) -> bool { matches!(self, ExtExpr::Lit(v) if v.is_zero()) } }
impl One for ExtExpr { fn one() -> Self { ExtExpr::Lit(SecureField::one()) } }
impl FieldExpOps for BaseExpr {
    fn inverse(&self) -> Self { match self { BaseExpr::Lit(v) => BaseExpr::Lit(v.inverse()), _ => panic!("an AIR constraint cannot invert a trace value (degree would be unbounded)") } }
}
impl FieldExpOps for ExtExpr {
    fn inverse(&self) -> Self { match self { ExtExpr::Lit(v) => ExtExpr::Lit(v.inverse()), _ => panic!("an AIR constraint cannot invert a trace value") } }
}

/// The recording evaluator for ONE component.  Components of a machine share the column allocator and the parameter table
/// (`Recorder::new(.., alloc.clone(), params.clone())`), exactly as they share `TraceLocationAllocator` upstream.
pub struct Recorder {
    g: G,
    log_size: u32,
    log_expand: u32,
    pending_fracs: Vec<(ExtExpr, ExtExpr)>,
}

impl Recorder {
    /// `log_expand` = `max_constraint_log_degree_bound() - log_size()` of the `FrameworkEval` (LOG_CONSTRAINT_DEGREE = 2 for the
    /// v1 main component, prover/src/components/mod.rs:12,44-46)
    pub fn new(log_size: u32, log_expand: u32, alloc: Rc<RefCell<Allocator>>, params: Rc<RefCell<ParamTable>>) -> Self {
        let g = Graph { nodes: vec![], cse: HashMap::new(), masks: vec![], mask_slot: HashMap::new(), constraints: vec![], fracs: vec![], batching: None,
                        cumsum_shift_param: NO_PARAM, interaction_col0: None, alloc, params };
        Recorder { g: Rc::new(RefCell::new(g)), log_size, log_expand, pending_fracs: vec![] }
    }
    fn mask_node(&self, tree: u32, col: u32, off: i32) -> BaseExpr {
        let id = { let mut g = self.g.borrow_mut(); let s = g.slot(tree, col, off); g.node(Op::Mask, s, 0, Kind::B) };
        BaseExpr::Node(GraphRef(self.g.clone()), id)
    }
    fn take_col(&self, tree: usize) -> u32 {
        let mut g = self.g.borrow_mut();
        let col = { let mut a = g.alloc.borrow_mut(); let c = a.next_col[tree]; a.next_col[tree] += 1; c };
        if tree == INTERACTION_TRACE_IDX && g.interaction_col0.is_none() { g.interaction_col0 = Some(col); }
        col
    }
    /// lookup-element handles of a relation as EF values (what `Relation::combine` multiplies with): (z, [alpha^i])
    pub fn relation_elements(&self, name: &str, size: usize) -> (ExtExpr, Vec<ExtExpr>) {
        let rp = self.g.borrow().params.borrow_mut().relation(name, size);
        let mk = |idx: u32| { let id = self.g.borrow_mut().node(Op::Param, idx, 0, Kind::E); ExtExpr::Node(GraphRef(self.g.clone()), id) };
        (mk(rp.z), rp.alpha_powers.iter().map(|&i| mk(i)).collect())
    }
    fn frac_sum(fr: &[(ExtExpr, ExtExpr)]) -> (ExtExpr, ExtExpr) {
        let (mut num, mut den) = fr[0].clone();
        for (n2, d2) in &fr[1..] { let nn = d2.clone() * num.clone() + den.clone() * n2.clone(); den = den * d2.clone(); num = nn; }
        (num, den)
    }
    /// number of secure LogUp columns this component adds to the interaction tree (`finalize_logup*`: one per batch) — 4 base-field
    /// coordinate columns each
    pub fn n_logup_columns(&self) -> usize {
        self.g.borrow().batching.as_ref().map_or(0, |b| b.iter().copied().max().map_or(0, |m| m + 1))
    }
    /// Serialise this component (call after `FrameworkEval::evaluate(recorder)` returned it).
    pub fn finish(self) -> Vec<u32> {
        assert!(self.pending_fracs.is_empty() || self.g.borrow().batching.is_some(), "logup fractions were added but finalize_logup was not called");
        let g = self.g.borrow();
        let sinks: Vec<(u32, u32, Option<u32>)> = g.constraints.iter().map(|&n| (if g.nodes[n as usize].kind == Kind::B { OP_CONSTRB } else { OP_CONSTRE }, n, None)).collect();
        let (prog, nb, ne) = emit(&g.nodes, &sinks);
        let lsinks: Vec<(u32, u32, Option<u32>)> = g.fracs.iter().map(|&(n, d)| (OP_FRAC, n, Some(d))).collect();
        let (lprog, lb, le) = if lsinks.is_empty() { (vec![], 0, 0) } else { emit(&g.nodes, &lsinks) };
        let mut w = vec![self.log_size, self.log_expand, g.constraints.len() as u32, g.masks.len() as u32];
        for &(t, c, o) in &g.masks { w.extend_from_slice(&[t, c, o as u32]); }
        w.extend_from_slice(&[nb, ne, prog.len() as u32]);
        for ins in &prog { w.extend_from_slice(ins); }
        w.extend_from_slice(&[g.fracs.len() as u32, lb, le, lprog.len() as u32]);
        for ins in &lprog { w.extend_from_slice(ins); }
        if let Some(b) = &g.batching { w.extend(b.iter().map(|&x| x as u32)); }
        w.extend_from_slice(&[g.cumsum_shift_param, g.interaction_col0.unwrap_or(0)]);
        w
    }
}

/// All components of a machine -> the word stream of `nb200_air_load` ('NBAR', version 1, n_params, n_components, ...)
pub struct AirBytecode;
impl AirBytecode {
    pub fn assemble(n_params: u32, components: Vec<Vec<u32>>) -> Vec<u32> {
        let mut w = vec![0x5241_424E, 1, n_params, components.len() as u32];
        for c in components { w.extend(c); }
        w
    }
}

impl EvalAtRow for Recorder {
    type F = BaseExpr;
    type EF = ExtExpr;

    fn next_interaction_mask<const N: usize>(&mut self, interaction: usize, offsets: [isize; N]) -> [Self::F; N] {
        let col = self.take_col(interaction);
        offsets.map(|o| self.mask_node(interaction as u32, col, o as i32))
    }
    fn get_preprocessed_column(&mut self, column: PreProcessedColumnId) -> Self::F {
        // preprocessed columns are addressed by id: the first request allocates the next column of tree 0 (air.py:preprocessed_column)
        let col = { let g = self.g.borrow(); let mut a = g.alloc.borrow_mut();
                    if let Some(&c) = a.preprocessed.get(&column.id) { c } else { let c = a.next_col[PREPROCESSED_TRACE_IDX]; a.next_col[PREPROCESSED_TRACE_IDX] += 1; a.preprocessed.insert(column.id.clone(), c); c } };
        self.mask_node(PREPROCESSED_TRACE_IDX as u32, col, 0)
    }
    fn next_extension_interaction_mask<const N: usize>(&mut self, interaction: usize, offsets: [isize; N]) -> [Self::EF; N] {
        let cols: Vec<u32> = (0..4).map(|_| self.take_col(interaction)).collect();
        offsets.map(|o| {
            let mut g = self.g.borrow_mut();
            let slots: Vec<u32> = cols.iter().map(|&c| g.slot(interaction as u32, c, o as i32)).collect();
            assert!(slots.windows(2).all(|w| w[1] == w[0] + 1), "the 4 coordinate slots of an extension mask must be consecutive (OP_LOADME)");
            let id = g.node(Op::MaskE, slots[0], 0, Kind::E);
            ExtExpr::Node(GraphRef(self.g.clone()), id)
        })
    }
    fn add_constraint<Gc>(&mut self, constraint: Gc) where Self::EF: Mul<Gc, Output = Self::EF> + From<Gc> {
        // upstream multiplies by the random coefficient here; the library applies the coefficient powers itself, so only the value is
        // recorded.  A base-field constraint stays a base-field sink when `Gc = F` (From<F> lifts through btoe, which `emit` sees).
        let e: ExtExpr = constraint.into();
        let id = match &e { ExtExpr::Node(_, id) => { let g = self.g.borrow(); let n = g.nodes[*id as usize]; if n.op == Op::BtoE { n.a } else { *id } }
                            ExtExpr::Lit(_) => e.id_in(&self.g) };
        self.g.borrow_mut().constraints.push(id);
    }
    fn combine_ef(values: [Self::F; 4]) -> Self::EF {
        // c0 + c1 i + c2 u + c3 iu with literal basis elements
        let basis = [SecureField::from_m31_array([1, 0, 0, 0].map(BaseField::from_u32_unchecked)), SecureField::from_m31_array([0, 1, 0, 0].map(BaseField::from_u32_unchecked)),
                     SecureField::from_m31_array([0, 0, 1, 0].map(BaseField::from_u32_unchecked)), SecureField::from_m31_array([0, 0, 0, 1].map(BaseField::from_u32_unchecked))];
        let mut acc = ExtExpr::zero();
        for (v, b) in values.into_iter().zip(basis) { acc = acc + v * b; }
        acc
    }
    fn add_to_relation<R: Relation<Self::F, Self::EF>>(&mut self, entry: RelationEntry<'_, Self::F, Self::EF, R>) {
        // Fraction::new(multiplicity, relation.combine(values)) — the relation's lookup elements must have been created through
        // `relation_elements` so that `combine` yields parameter nodes (see prover-patch/src/cuda/lookups.rs)
        let den = entry.relation.combine(entry.values);
        self.write_logup_frac(Fraction::new(entry.multiplicity.clone(), den));
    }
    fn write_logup_frac(&mut self, fraction: Fraction<Self::EF, Self::EF>) {
        let (n, d) = (fraction.numerator.id_in(&self.g), fraction.denominator.id_in(&self.g));
        self.g.borrow_mut().fracs.push((n, d));
        self.pending_fracs.push((fraction.numerator, fraction.denominator));
    }
    fn finalize_logup_batched(&mut self, batching: &Batching) {
        assert!(self.g.borrow().batching.is_none() && batching.len() == self.pending_fracs.len() && !batching.is_empty());
        let last = *batching.iter().max().unwrap();
        let shift = { let g = self.g.borrow(); let mut p = g.params.borrow_mut(); p.alloc("cumsum_shift") };
        { let mut g = self.g.borrow_mut(); g.batching = Some(batching.clone()); g.cumsum_shift_param = shift; }
        let fr = self.pending_fracs.clone();
        let of_batch = |b: usize| -> Vec<(ExtExpr, ExtExpr)> { batching.iter().zip(&fr).filter(|(x, _)| **x == b).map(|(_, f)| f.clone()).collect() };
        let mut prev_col: Option<ExtExpr> = None;
        for b in 0..last {
            let (num, den) = Self::frac_sum(&of_batch(b));
            let [cur] = self.next_extension_interaction_mask(INTERACTION_TRACE_IDX, [0]);
            let diff = match &prev_col { None => cur.clone(), Some(p) => cur.clone() - p.clone() };
            prev_col = Some(cur);
            self.add_constraint(diff * den - num);
        }
        let (num, den) = Self::frac_sum(&of_batch(last));
        let [prev_row, cur] = self.next_extension_interaction_mask(INTERACTION_TRACE_IDX, [-1, 0]);
        let mut diff = cur - prev_row;
        if let Some(p) = prev_col { diff = diff - p; }
        let shift_e = { let id = self.g.borrow_mut().node(Op::Param, shift, 0, Kind::E); ExtExpr::Node(GraphRef(self.g.clone()), id) };
        self.add_constraint((diff + shift_e) * den - num);
    }
    fn finalize_logup(&mut self) { let n = self.pending_fracs.len(); self.finalize_logup_batched(&(0..n).collect()) }
    fn finalize_logup_in_pairs(&mut self) { let n = self.pending_fracs.len(); self.finalize_logup_batched(&(0..n).map(|k| k / 2).collect()) }
}

fn operands(n: &Node) -> Vec<u32> {
    match n.op { Op::Add | Op::Sub | Op::Mul | Op::AddEb | Op::SubEb | Op::MulEb => vec![n.a, n.b], Op::Neg | Op::BtoE => vec![n.a], _ => vec![] }
}
fn is_leaf(n: &Node) -> bool { matches!(n.op, Op::Mask | Op::MaskE | Op::Const | Op::Param) }

/// air.py `_emit`: code for the sinks in declaration order — before each sink the not yet computed part of its DAG in post-order
/// (leaves re-materialised per sink), then the sink; virtual registers are reused once an instance's last consumer is emitted.
fn emit(nodes: &[Node], sinks: &[(u32, u32, Option<u32>)]) -> (Vec<[u32; 4]>, u32, u32) {
    #[derive(Clone, Copy)] enum Item { N(u32), S(usize) }
    let mut seq: Vec<Item> = vec![];
    let mut done: std::collections::HashSet<u32> = Default::default();
    for (idx, s) in sinks.iter().enumerate() {
        let mut fresh: std::collections::HashSet<u32> = Default::default();
        for root in [Some(s.1), s.2].into_iter().flatten() {
            if done.contains(&root) || fresh.contains(&root) { continue; }
            let mut stack = vec![(root, false)];
            while let Some((n, expanded)) = stack.pop() {
                if done.contains(&n) || fresh.contains(&n) { continue; }
                if expanded { if is_leaf(&nodes[n as usize]) { fresh.insert(n); } else { done.insert(n); } seq.push(Item::N(n)); continue; }
                stack.push((n, true));
                for x in operands(&nodes[n as usize]).into_iter().rev() { if !done.contains(&x) && !fresh.contains(&x) { stack.push((x, false)); } }
            }
        }
        seq.push(Item::S(idx));
    }
    // instance liveness: backward scan
    let mut last: HashMap<u32, usize> = HashMap::new();
    let mut inst_last: HashMap<usize, Option<usize>> = HashMap::new();
    for pos in (0..seq.len()).rev() {
        match seq[pos] {
            Item::N(v) => { inst_last.insert(pos, last.remove(&v)); for x in operands(&nodes[v as usize]) { last.entry(x).or_insert(pos); } }
            Item::S(i) => { for x in [Some(sinks[i].1), sinks[i].2].into_iter().flatten() { last.entry(x).or_insert(pos); } }
        }
    }
    let mut free_at: HashMap<usize, Vec<usize>> = HashMap::new();
    for (&d, &l) in &inst_last { if let Some(l) = l { free_at.entry(l).or_default().push(d); } }
    for v in free_at.values_mut() { v.sort(); }   // deterministic register reuse (air.py iterates a dict in insertion order = ascending definition position)
    let (mut reg, mut reg_of_def): (HashMap<u32, u32>, HashMap<usize, u32>) = Default::default();
    let mut free: [Vec<u32>; 2] = [vec![], vec![]];
    let mut nreg = [0u32; 2];
    let mut out: Vec<[u32; 4]> = vec![];
    let ki = |k: Kind| if k == Kind::B { 0 } else { 1 };
    for (pos, item) in seq.iter().enumerate() {
        let mut pending: Option<(Node, u32, u32)> = None;
        match *item {
            Item::S(i) => { let (sop, n1, n2) = sinks[i]; out.push([sop, 0, reg[&n1], n2.map(|n| reg[&n]).unwrap_or(0)]); }
            Item::N(v) => { let n = nodes[v as usize]; let ops = operands(&n);
                            pending = Some((n, ops.first().map(|a| reg[a]).unwrap_or(0), ops.get(1).map(|b| reg[b]).unwrap_or(0))); }
        }
        if let Some(ds) = free_at.get(&pos) { for &d in ds { if let Item::N(v) = seq[d] { free[ki(nodes[v as usize].kind)].push(reg_of_def[&d]); } } }
        let Some((n, ra, rb)) = pending else { continue };
        let Item::N(v) = *item else { unreachable!() };
        let k = ki(n.kind);
        let r = free[k].pop().unwrap_or_else(|| { nreg[k] += 1; nreg[k] - 1 });
        reg.insert(v, r); reg_of_def.insert(pos, r);
        out.push(match n.op {
            Op::Mask => [OP_LOADM, r, n.a, 0], Op::MaskE => [OP_LOADME, r, n.a, 0], Op::Const => [OP_CONSTB, r, n.a, 0], Op::Param => [OP_PARAME, r, n.a, 0],
            Op::BtoE => [OP_BTOE, r, ra, 0],
            Op::Neg => [if n.kind == Kind::B { OP_NEGB } else { OP_NEGE }, r, ra, 0],
            Op::Add => [if n.kind == Kind::B { OP_ADDB } else { OP_ADDE }, r, ra, rb],
            Op::Sub => [if n.kind == Kind::B { OP_SUBB } else { OP_SUBE }, r, ra, rb],
            Op::Mul => [if n.kind == Kind::B { OP_MULB } else { OP_MULE }, r, ra, rb],
            Op::AddEb => [OP_ADDEB, r, ra, rb], Op::SubEb => [OP_SUBEB, r, ra, rb], Op::MulEb => [OP_MULEB, r, ra, rb],
        });
        if inst_last[&pos].is_none() { free[k].push(r); }
    }
    (out, nreg[0], nreg[1])
}
