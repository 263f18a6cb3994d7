//! taiga_halo2/tests/gpu_proofs_verify.rs -- closes the parity gap of DESIGN.md section 3.
//! Build taiga_halo2 against the patched halo2_proofs (feature "gpu") on a machine with cargo and a B200, then
//!     cargo test --release --features gpu gpu_proofs_verify
//! Every proof below is MADE by libtaiga_b200.so (through the unchanged `Proof::create`) and CHECKED by the reference's own
//! CPU verifier (`Proof::verify`, taiga_halo2/src/proof.rs:45-54 -> halo2_proofs::plonk::verify_proof with SingleVerifier).
//! Not compiled in the build image (no cargo).
use halo2_proofs::dev::MockProver;
use rand::rngs::OsRng;
use taiga_halo2::{
    circuit::resource_logic_circuit::ResourceLogicVerifyingInfoTrait,
    circuit::resource_logic_examples::TrivialResourceLogicCircuit,
    compliance::tests::random_compliance_info,
    constant::{
        COMPLIANCE_CIRCUIT_PARAMS_SIZE, COMPLIANCE_PROVING_KEY, COMPLIANCE_VERIFYING_KEY, SETUP_PARAMS_MAP,
    },
    proof::Proof,
    resource::tests::random_resource,
    shielded_ptx::testing::create_shielded_ptx,
    transaction::{ShieldedPartialTxBundle, Transaction, TransparentPartialTxBundle},
};

/// compliance_circuit.rs:330-374 with the GPU behind `Proof::create`.
#[test]
fn gpu_proofs_verify_compliance() {
    let mut rng = OsRng;
    let info = random_compliance_info(&mut rng); // compliance.rs:244, the fixture of compliance_circuit.rs:330-374
    let (instance, circuit) = info.build();
    let instance = instance.to_instance();
    // the witness satisfies the circuit (otherwise a rejection below would say nothing about the prover)
    assert_eq!(MockProver::run(COMPLIANCE_CIRCUIT_PARAMS_SIZE, &circuit, vec![instance.clone()]).unwrap().verify(), Ok(()));
    let params = SETUP_PARAMS_MAP.get(&COMPLIANCE_CIRCUIT_PARAMS_SIZE).unwrap();
    let proof = Proof::create(&COMPLIANCE_PROVING_KEY, params, circuit, &[&instance], &mut rng).unwrap();
    assert_eq!(proof.inner().len(), 4480); // 4676 = 4 + 4480 + 6 * 32, taiga_api.rs:109
    proof.verify(&COMPLIANCE_VERIFYING_KEY, params, &[&instance]).expect("stock verifier rejected a GPU proof");
    // and it is a proof OF THIS INSTANCE: a different public input must be rejected
    let mut wrong = instance.clone();
    wrong[0] += pasta_curves::pallas::Base::one();
    assert!(proof.verify(&COMPLIANCE_VERIFYING_KEY, params, &[&wrong]).is_err());
}

/// resource_logic_examples.rs:156-174 + get_verifying_info (resource_logic_examples.rs:117-134): TrivialVP.
#[test]
fn gpu_proofs_verify_trivial_vp() {
    let mut rng = OsRng;
    let circuit = TrivialResourceLogicCircuit::new(
        random_resource(&mut rng).get_nf().unwrap().inner(),
        [(); 2].map(|_| random_resource(&mut rng)),
        [(); 2].map(|_| random_resource(&mut rng)),
    );
    let info = circuit.get_verifying_info(); // keygen (cached pk) + Proof::create through the GPU
    info.verify().expect("stock verifier rejected a GPU TrivialVP proof"); // ResourceLogicVerifyingInfo::verify, resource_logic_circuit.rs:155-161
}

/// taiga_api.rs:254-352 / transaction.rs:350-374: a whole shielded partial transaction (2 Compliance + the VP proofs),
/// borsh round trip and `Transaction::execute` (= verify every proof + the binding signature).
#[test]
fn gpu_proofs_verify_ptx_and_transaction() {
    let mut rng = OsRng;
    let (ptx, r) = create_shielded_ptx(); // shielded_ptx.rs:432-619: every proof inside comes from the GPU
    ptx.clone().execute().expect("stock verifier rejected a GPU-built partial transaction");
    let bytes = borsh::to_vec(&ptx).unwrap();
    let back: taiga_halo2::shielded_ptx::ShieldedPartialTransaction = borsh::BorshDeserialize::deserialize(&mut bytes.as_ref()).unwrap();
    back.execute().unwrap();
    let tx = Transaction::build(&mut rng, ShieldedPartialTxBundle::new(vec![ptx]), TransparentPartialTxBundle::default(), vec![r]).unwrap();
    tx.execute().expect("transaction with GPU proofs must execute");
}
