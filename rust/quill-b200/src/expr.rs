//! `VirtualPolyExpr` (hyperplonk/src/utils/virtual_polynomial.rs:9-18) -> the flat node array of the C ABI.
use ark_bn254::Fr;
use quill_b200_sys::{qz_expr_node, QZ_EX_ADD, QZ_EX_CONST, QZ_EX_INPUT, QZ_EX_MUL};
use quill_hyperplonk::utils::virtual_polynomial::VirtualPolyExpr;

/// Post-order walk: children before parents, the root last; constants are collected in order of first appearance.
pub fn flatten(e: &VirtualPolyExpr<Fr>) -> (Vec<qz_expr_node>, Vec<Fr>) {
    fn go(e: &VirtualPolyExpr<Fr>, nodes: &mut Vec<qz_expr_node>, consts: &mut Vec<Fr>) -> u32 {
        let node = match e {
            VirtualPolyExpr::Input(i) => qz_expr_node { op: QZ_EX_INPUT, a: *i as u32, b: 0 },
            VirtualPolyExpr::Const(c) => {
                consts.push(*c);
                qz_expr_node { op: QZ_EX_CONST, a: (consts.len() - 1) as u32, b: 0 }
            }
            VirtualPolyExpr::Add(l, r) => {
                let (a, b) = (go(l, nodes, consts), go(r, nodes, consts));
                qz_expr_node { op: QZ_EX_ADD, a, b }
            }
            VirtualPolyExpr::Mul(l, r) => {
                let (a, b) = (go(l, nodes, consts), go(r, nodes, consts));
                qz_expr_node { op: QZ_EX_MUL, a, b }
            }
        };
        nodes.push(node);
        (nodes.len() - 1) as u32
    }
    let (mut nodes, mut consts) = (Vec::new(), Vec::new());
    go(e, &mut nodes, &mut consts);
    (nodes, consts)
}

/// Two expressions over ONE constants array (qz_logup_denominators takes h and m that way): the second expression's
/// Const indices are shifted past the first's.
pub fn flatten_pair(h: &VirtualPolyExpr<Fr>, m: &VirtualPolyExpr<Fr>) -> (Vec<qz_expr_node>, Vec<qz_expr_node>, Vec<Fr>) {
    let (nh, mut consts) = flatten(h);
    let (mut nm, cm) = flatten(m);
    let shift = consts.len() as u32;
    for n in nm.iter_mut() {
        if n.op == QZ_EX_CONST {
            n.a += shift;
        }
    }
    consts.extend(cm);
    (nh, nm, consts)
}
